// host_expand.h -- column indices from visibility words on the host (see host_expand.cpp)
#pragma once
#include <stdint.h>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

namespace fluxb200 {

// positions of the set bits of words[0..nwords), ascending, as int32 or int64; returns how many.
// row_entries = popcount of the words if the caller knows it (-1: counted here); exactly that many
// entries of `out` are written, never more.
int64_t expand_words(const uint32_t *words, int nwords, int index_width, void *out, int64_t row_entries = -1);

int64_t count_bits(const uint32_t *words, int nwords);

// the rows of one sub-slab: words (mr x nwords, host) -> indices[offs[r] .. offs[r+1])
struct ExpandTask {
    const uint32_t *words = nullptr;
    int nwords = 0;
    size_t mr = 0;
    std::vector<int64_t> offs; // mr + 1 positions in `indices`
    void *indices = nullptr;
    int index_width = 4;
    // optional second job of the same sub-slab: values that were copied out into a page-locked staging
    // slot go on to the caller's ordinary (pageable) array; every piece copies its share of the bytes
    const char *copy_src = nullptr;
    char *copy_dst = nullptr;
    size_t copy_bytes = 0;
    int pieces = 1;               // row ranges handed to the workers
    std::atomic<int> pending{0};  // pieces not finished yet
    std::atomic<int> queued{0};   // set once the stream callback has submitted it
    std::atomic<int> mismatch{0}; // a row whose bit count differs from its CSR row length
    class HostExpander *owner = nullptr;
};

class HostExpander {
public:
    ~HostExpander();
    void start(int nthreads);
    int threads() const { return (int)workers_.size(); }
    void submit(ExpandTask *t); // callable from a CUDA host-function callback (no CUDA calls inside)
    void wait(ExpandTask *t);

private:
    void run();
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    std::deque<std::pair<ExpandTask *, int>> queue_;
    bool stop_ = false;
};

} // namespace fluxb200
