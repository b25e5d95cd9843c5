// host_api.cu — host-buffer entry point (evogp_SR_fitness_host): the whole
// "forest on the host -> fitness on the host" trip a non-torch caller makes.
//
// The trip is PCIe-bound (the evaluation of 100000 trees takes 0.25 ms, their fixed-width rows are 51 MB), so the
// job of this file is to move fewer bytes and to keep the link busy:
//   * only what the evaluator reads crosses PCIe: the VALID PREFIX of node_value and node_type of every tree, packed
//     back to back (6 B per node, mean prefix 26 of 64 slots at BASELINE configs[1]: 16 MB instead of 51 MB), plus one
//     32-bit offset per tree.  subtree_size never travels: the lowering pass rebuilds sizes from the arities;
//   * the packing runs on a small pool of host threads (the rows are gathered into pinned staging buffers);
//   * the forest is cut into row chunks that alternate between two streams and two staging buffers, so packing chunk
//     c + 1, the H2D copy of chunk c + 1 and lowering + replay of chunk c overlap.
// Staging buffers and worker threads are created once and cached across calls.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include "common.cuh"

int evogp_sr_fitness_packed(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen, unsigned outLen, int useMSE,
                            const float *value, const int16_t *type, const unsigned *offsets, const float *variables,
                            const float *labels, float *fitnesses, void *workspace, size_t workspace_bytes, void *stream);

namespace evogp {
namespace {

// ---- a minimal fork-join pool: run(f) calls f(worker, nworkers) on every worker and returns when all are done ----
class WorkerPool {
public:
    explicit WorkerPool(int n) : n_(n) {
        for (int i = 0; i < n_; ++i) threads_.emplace_back([this, i] { loop(i); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++generation_;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    int size() const { return n_; }
    void start(const std::function<void(int, int)> &f) {      // f must stay alive until wait() returns
        std::lock_guard<std::mutex> lk(mu_);
        job_ = &f;
        pending_ = n_;
        ++generation_;
        cv_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int id) {
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(int, int)> *job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
                job = job_;
            }
            (*job)(id, n_);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)> *job_ = nullptr;
    unsigned long generation_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

struct Staging {
    int device = -1;
    size_t chunk_rows = 0, L = 0;
    // pinned host staging: the packed prefixes of the WHOLE population (chunk after chunk) and per-chunk offsets;
    // device images of one chunk, double-buffered
    size_t pop_cap = 0;
    float *h_value = nullptr;
    int16_t *h_type = nullptr;
    unsigned *h_off = nullptr;       // per chunk: nr + 1 offsets relative to the chunk's first node
    uint16_t *h_len = nullptr;
    float *value[2] = {nullptr, nullptr};
    int16_t *type[2] = {nullptr, nullptr};
    unsigned *off[2] = {nullptr, nullptr};
    void *ws[2] = {nullptr, nullptr};
    size_t ws_bytes = 0;
    float *fitness[2] = {nullptr, nullptr};
    float *X = nullptr, *labels = nullptr;
    size_t x_cap = 0, lab_cap = 0;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t data_ready = nullptr;
    WorkerPool *pool = nullptr;
};
Staging g_st;
std::mutex g_mu;

void release_locked() {
    if (g_st.device >= 0) {
        cudaSetDevice(g_st.device);
        if (g_st.h_value) cudaFreeHost(g_st.h_value);
        if (g_st.h_type) cudaFreeHost(g_st.h_type);
        if (g_st.h_off) cudaFreeHost(g_st.h_off);
        free(g_st.h_len);
        for (int i = 0; i < 2; ++i) {
            cudaFree(g_st.value[i]); cudaFree(g_st.type[i]); cudaFree(g_st.off[i]);
            cudaFree(g_st.ws[i]); cudaFree(g_st.fitness[i]);
            if (g_st.stream[i]) cudaStreamDestroy(g_st.stream[i]);
        }
        cudaFree(g_st.X); cudaFree(g_st.labels);
        if (g_st.data_ready) cudaEventDestroy(g_st.data_ready);
    }
    delete g_st.pool;
    g_st = Staging();
}

int prepare(int device, size_t pop, size_t rows, size_t L, size_t xbytes, size_t lbytes) {
    if (g_st.device != device || g_st.chunk_rows < rows || g_st.L != L || g_st.pop_cap < pop) {
        WorkerPool *pool = g_st.pool;      // threads survive a re-size of the buffers
        g_st.pool = nullptr;
        release_locked();
        g_st.pool = pool;
        EVOGP_CUDA(cudaSetDevice(device));
        g_st.device = device;
        g_st.chunk_rows = rows;
        g_st.L = L;
        g_st.ws_bytes = evogp_eval_workspace_bytes((unsigned)rows, (unsigned)L);
        g_st.pop_cap = pop;
        const size_t chunks = (pop + rows - 1) / rows;
        EVOGP_CUDA(cudaHostAlloc(&g_st.h_value, pop * L * sizeof(float), cudaHostAllocDefault));
        EVOGP_CUDA(cudaHostAlloc(&g_st.h_type, pop * L * sizeof(int16_t), cudaHostAllocDefault));
        EVOGP_CUDA(cudaHostAlloc(&g_st.h_off, (pop + chunks) * sizeof(unsigned), cudaHostAllocDefault));
        g_st.h_len = static_cast<uint16_t *>(malloc(pop * sizeof(uint16_t)));
        for (int i = 0; i < 2; ++i) {
            EVOGP_CUDA(cudaMalloc(&g_st.value[i], rows * L * sizeof(float)));
            EVOGP_CUDA(cudaMalloc(&g_st.type[i], rows * L * sizeof(int16_t)));
            EVOGP_CUDA(cudaMalloc(&g_st.off[i], (rows + 1) * sizeof(unsigned)));
            EVOGP_CUDA(cudaMalloc(&g_st.ws[i], g_st.ws_bytes));
            EVOGP_CUDA(cudaMalloc(&g_st.fitness[i], rows * sizeof(float)));
            EVOGP_CUDA(cudaStreamCreateWithFlags(&g_st.stream[i], cudaStreamNonBlocking));
        }
        EVOGP_CUDA(cudaEventCreateWithFlags(&g_st.data_ready, cudaEventDisableTiming));
    }
    if (!g_st.pool) {
        // EVOGP_HOST_THREADS: packing threads (default: a quarter of the host's hardware threads, 2..16)
        const char *e = getenv("EVOGP_HOST_THREADS");
        int n = e ? atoi(e) : (int)std::thread::hardware_concurrency() / 4;
        n = std::max(2, std::min(n, 16));
        if (e && atoi(e) == 1) n = 1;
        g_st.pool = new WorkerPool(n);
    }
    if (g_st.x_cap < xbytes) {
        cudaFree(g_st.X);
        EVOGP_CUDA(cudaMalloc(&g_st.X, xbytes));
        g_st.x_cap = xbytes;
    }
    if (g_st.lab_cap < lbytes) {
        cudaFree(g_st.labels);
        EVOGP_CUDA(cudaMalloc(&g_st.labels, lbytes));
        g_st.lab_cap = lbytes;
    }
    return EVOGP_OK;
}

}  // namespace
}  // namespace evogp

using namespace evogp;

extern "C" void evogp_host_release(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    release_locked();
}

extern "C" int evogp_SR_fitness_host(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen,
                                     unsigned outLen, int useMSE, const float *value, const int16_t *type,
                                     const int16_t *subtree_size, const float *variables, const float *labels,
                                     float *fitnesses, int device) {
    EVOGP_REQUIRE(popSize > 0 && dataPoints > 0 && gpLen > 0 && varLen > 0 && outLen > 0, "empty problem");
    std::lock_guard<std::mutex> lk(g_mu);
    // chunk so that a few chunks are in flight even for small populations, capped at 64 Ki rows
    // (EVOGP_HOST_CHUNKS overrides the default of 8 chunks; tuning knob)
    static const int n_chunks = [] { const char *e = getenv("EVOGP_HOST_CHUNKS"); int v = e ? atoi(e) : 8; return v < 1 ? 1 : v; }();
    size_t rows = (popSize + n_chunks - 1) / n_chunks;
    if (rows < 4096) rows = popSize < 4096 ? popSize : 4096;
    if (rows > 65536) rows = 65536;
    const size_t L = gpLen;
    const size_t xbytes = (size_t)dataPoints * varLen * sizeof(float), lbytes = (size_t)dataPoints * outLen * sizeof(float);
    int rc = prepare(device, popSize, rows, L, xbytes, lbytes);
    if (rc) return rc;
    EVOGP_CUDA(cudaSetDevice(device));
    Staging &s = g_st;
    EVOGP_CUDA(cudaMemcpyAsync(s.X, variables, xbytes, cudaMemcpyHostToDevice, s.stream[0]));
    EVOGP_CUDA(cudaMemcpyAsync(s.labels, labels, lbytes, cudaMemcpyHostToDevice, s.stream[0]));
    EVOGP_CUDA(cudaEventRecord(s.data_ready, s.stream[0]));
    EVOGP_CUDA(cudaStreamWaitEvent(s.stream[1], s.data_ready, 0));
    const size_t chunks = (popSize + rows - 1) / rows;
    WorkerPool &pool = *s.pool;
    static const bool trace = getenv("EVOGP_HOST_TRACE") != nullptr;    // phase times on stderr (developer aid)
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    // ---- 1. lengths: the only column of subtree_size the evaluator needs (one cache line per tree: spread over the pool) ----
    uint16_t *len = s.h_len;
    const std::function<void(int, int)> lens_job = [&](int w, int nw) {
        const size_t lo = (size_t)popSize * w / nw, hi = (size_t)popSize * (w + 1) / nw;
        for (size_t r = lo; r < hi; ++r) {
            const int v = subtree_size[r * L];
            len[r] = (uint16_t)(v < 0 ? 0 : (v > (int)L ? (int)L : v));   // impossible lengths become empty rows (-> NaN fitness)
        }
    };
    pool.start(lens_job);
    pool.wait();
    const double t_lens = since(t_begin);
    // ---- 2. offsets per chunk (relative to the chunk's first node) and the chunk's place in the staging buffers ----
    std::vector<size_t> node0(chunks + 1, 0);
    for (size_t c = 0, r0 = 0; c < chunks; ++c, r0 += rows) {
        const size_t nr = std::min(rows, (size_t)popSize - r0);
        unsigned *off = s.h_off + r0 + c;
        unsigned total = 0;
        for (size_t r = 0; r < nr; ++r) {
            off[r] = total;
            total += len[r0 + r];
        }
        off[nr] = total;
        node0[c + 1] = node0[c] + total;
    }
    // ---- 3. pack chunk after chunk on the pool; this thread enqueues each chunk as soon as it is packed ----
    std::vector<std::atomic<int>> packed(chunks);
    for (auto &p : packed) p.store(0, std::memory_order_relaxed);
    const std::function<void(int, int)> pack_job = [&](int w, int nw) {
        for (size_t c = 0, r0 = 0; c < chunks; ++c, r0 += rows) {
            const size_t nr = std::min(rows, (size_t)popSize - r0);
            const unsigned *off = s.h_off + r0 + c;
            float *hv = s.h_value + node0[c];
            int16_t *ht = s.h_type + node0[c];
            const size_t lo = nr * (size_t)w / nw, hi = nr * (size_t)(w + 1) / nw;
            for (size_t r = lo; r < hi; ++r) {
                const unsigned n = off[r + 1] - off[r];
                std::memcpy(hv + off[r], value + (r0 + r) * L, n * sizeof(float));
                std::memcpy(ht + off[r], type + (r0 + r) * L, n * sizeof(int16_t));
            }
            packed[c].fetch_add(1, std::memory_order_release);
        }
    };
    const double t_offsets = since(t_begin);
    pool.start(pack_job);
    const int nw = pool.size();
    double t_first = 0, t_lastpack = 0;
    for (size_t c = 0, r0 = 0; c < chunks; ++c, r0 += rows) {
        const int b = (int)(c & 1);
        const size_t nr = std::min(rows, (size_t)popSize - r0);
        cudaStream_t st = s.stream[b];
        while (packed[c].load(std::memory_order_acquire) < nw) std::this_thread::yield();
        if (c == 0) t_first = since(t_begin);
        if (c + 1 == chunks) t_lastpack = since(t_begin);
        const size_t total = node0[c + 1] - node0[c];
        rc = EVOGP_OK;
        cudaError_t e = cudaMemcpyAsync(s.value[b], s.h_value + node0[c], total * sizeof(float), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s.type[b], s.h_type + node0[c], total * sizeof(int16_t), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s.off[b], s.h_off + r0 + c, (nr + 1) * sizeof(unsigned), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess)
            rc = evogp_sr_fitness_packed((unsigned)nr, dataPoints, gpLen, varLen, outLen, useMSE, s.value[b], s.type[b], s.off[b], s.X,
                                         s.labels, s.fitness[b], s.ws[b], s.ws_bytes, st);
        if (e == cudaSuccess && rc == EVOGP_OK)
            e = cudaMemcpyAsync(fitnesses + r0, s.fitness[b], nr * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess || rc != EVOGP_OK) {
            pool.wait();                                                   // the job references this frame
            if (e != cudaSuccess) {
                set_error("host path: %s", cudaGetErrorString(e));
                return EVOGP_ERR_CUDA;
            }
            return rc;
        }
    }
    pool.wait();
    const double t_enqueued = since(t_begin);
    EVOGP_CUDA(cudaStreamSynchronize(s.stream[0]));
    EVOGP_CUDA(cudaStreamSynchronize(s.stream[1]));
    if (trace)
        fprintf(stderr, "[evogp host] lens %.3f  offsets %.3f  first chunk packed %.3f  last chunk packed %.3f  all enqueued %.3f  done %.3f ms (%zu nodes, %d threads)\n",
                t_lens, t_offsets, t_first, t_lastpack, t_enqueued, since(t_begin), node0[chunks], nw);
    return EVOGP_OK;
}
