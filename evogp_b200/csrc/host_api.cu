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
    void run(const std::function<void(int, int)> &f) {
        std::unique_lock<std::mutex> lk(mu_);
        job_ = &f;
        pending_ = n_;
        ++generation_;
        cv_.notify_all();
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
    // pinned host staging (packed prefixes of one chunk) and their device images, double-buffered
    float *h_value[2] = {nullptr, nullptr};
    int16_t *h_type[2] = {nullptr, nullptr};
    unsigned *h_off[2] = {nullptr, nullptr};
    float *value[2] = {nullptr, nullptr};
    int16_t *type[2] = {nullptr, nullptr};
    unsigned *off[2] = {nullptr, nullptr};
    void *ws[2] = {nullptr, nullptr};
    size_t ws_bytes = 0;
    float *fitness[2] = {nullptr, nullptr};
    float *X = nullptr, *labels = nullptr;
    size_t x_cap = 0, lab_cap = 0;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t data_ready = nullptr, uploaded[2] = {nullptr, nullptr};
    WorkerPool *pool = nullptr;
};
Staging g_st;
std::mutex g_mu;

void release_locked() {
    if (g_st.device >= 0) {
        cudaSetDevice(g_st.device);
        for (int i = 0; i < 2; ++i) {
            if (g_st.h_value[i]) cudaFreeHost(g_st.h_value[i]);
            if (g_st.h_type[i]) cudaFreeHost(g_st.h_type[i]);
            if (g_st.h_off[i]) cudaFreeHost(g_st.h_off[i]);
            cudaFree(g_st.value[i]); cudaFree(g_st.type[i]); cudaFree(g_st.off[i]);
            cudaFree(g_st.ws[i]); cudaFree(g_st.fitness[i]);
            if (g_st.stream[i]) cudaStreamDestroy(g_st.stream[i]);
            if (g_st.uploaded[i]) cudaEventDestroy(g_st.uploaded[i]);
        }
        cudaFree(g_st.X); cudaFree(g_st.labels);
        if (g_st.data_ready) cudaEventDestroy(g_st.data_ready);
    }
    delete g_st.pool;
    g_st = Staging();
}

int prepare(int device, size_t rows, size_t L, size_t xbytes, size_t lbytes) {
    if (g_st.device != device || g_st.chunk_rows < rows || g_st.L != L) {
        WorkerPool *pool = g_st.pool;      // threads survive a re-size of the buffers
        g_st.pool = nullptr;
        release_locked();
        g_st.pool = pool;
        EVOGP_CUDA(cudaSetDevice(device));
        g_st.device = device;
        g_st.chunk_rows = rows;
        g_st.L = L;
        g_st.ws_bytes = evogp_eval_workspace_bytes((unsigned)rows, (unsigned)L);
        for (int i = 0; i < 2; ++i) {
            EVOGP_CUDA(cudaHostAlloc(&g_st.h_value[i], rows * L * sizeof(float), cudaHostAllocDefault));
            EVOGP_CUDA(cudaHostAlloc(&g_st.h_type[i], rows * L * sizeof(int16_t), cudaHostAllocDefault));
            EVOGP_CUDA(cudaHostAlloc(&g_st.h_off[i], (rows + 1) * sizeof(unsigned), cudaHostAllocDefault));
            EVOGP_CUDA(cudaMalloc(&g_st.value[i], rows * L * sizeof(float)));
            EVOGP_CUDA(cudaMalloc(&g_st.type[i], rows * L * sizeof(int16_t)));
            EVOGP_CUDA(cudaMalloc(&g_st.off[i], (rows + 1) * sizeof(unsigned)));
            EVOGP_CUDA(cudaMalloc(&g_st.ws[i], g_st.ws_bytes));
            EVOGP_CUDA(cudaMalloc(&g_st.fitness[i], rows * sizeof(float)));
            EVOGP_CUDA(cudaStreamCreateWithFlags(&g_st.stream[i], cudaStreamNonBlocking));
            EVOGP_CUDA(cudaEventCreateWithFlags(&g_st.uploaded[i], cudaEventDisableTiming));
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
    int rc = prepare(device, rows, L, xbytes, lbytes);
    if (rc) return rc;
    EVOGP_CUDA(cudaSetDevice(device));
    Staging &s = g_st;
    EVOGP_CUDA(cudaMemcpyAsync(s.X, variables, xbytes, cudaMemcpyHostToDevice, s.stream[0]));
    EVOGP_CUDA(cudaMemcpyAsync(s.labels, labels, lbytes, cudaMemcpyHostToDevice, s.stream[0]));
    EVOGP_CUDA(cudaEventRecord(s.data_ready, s.stream[0]));
    EVOGP_CUDA(cudaStreamWaitEvent(s.stream[1], s.data_ready, 0));
    int c = 0;
    for (size_t r0 = 0; r0 < popSize; r0 += rows, ++c) {
        const int b = c & 1;
        const size_t nr = (popSize - r0 < rows) ? popSize - r0 : rows;
        cudaStream_t st = s.stream[b];
        if (c >= 2) EVOGP_CUDA(cudaEventSynchronize(s.uploaded[b]));      // the copies out of this staging pair are done
        // ---- pack: offsets (serial, one column read per tree), then the prefixes in parallel ----
        unsigned *off = s.h_off[b];
        unsigned total = 0;
        for (size_t r = 0; r < nr; ++r) {
            int len = subtree_size[(r0 + r) * L];                          // the only column of subtree_size the evaluator needs
            len = len < 0 ? 0 : (len > (int)L ? (int)L : len);            // impossible lengths become empty rows (-> NaN fitness)
            off[r] = total;
            total += (unsigned)len;
        }
        off[nr] = total;
        float *hv = s.h_value[b];
        int16_t *ht = s.h_type[b];
        const std::function<void(int, int)> job = [&](int w, int nw) {
            const size_t lo = nr * (size_t)w / nw, hi = nr * (size_t)(w + 1) / nw;
            for (size_t r = lo; r < hi; ++r) {
                const unsigned len = off[r + 1] - off[r];
                std::memcpy(hv + off[r], value + (r0 + r) * L, len * sizeof(float));
                std::memcpy(ht + off[r], type + (r0 + r) * L, len * sizeof(int16_t));
            }
        };
        s.pool->run(job);
        // ---- upload, evaluate, download ----
        EVOGP_CUDA(cudaMemcpyAsync(s.value[b], hv, (size_t)total * sizeof(float), cudaMemcpyHostToDevice, st));
        EVOGP_CUDA(cudaMemcpyAsync(s.type[b], ht, (size_t)total * sizeof(int16_t), cudaMemcpyHostToDevice, st));
        EVOGP_CUDA(cudaMemcpyAsync(s.off[b], off, (nr + 1) * sizeof(unsigned), cudaMemcpyHostToDevice, st));
        EVOGP_CUDA(cudaEventRecord(s.uploaded[b], st));
        rc = evogp_sr_fitness_packed((unsigned)nr, dataPoints, gpLen, varLen, outLen, useMSE, s.value[b], s.type[b], s.off[b], s.X,
                                     s.labels, s.fitness[b], s.ws[b], s.ws_bytes, st);
        if (rc) return rc;
        EVOGP_CUDA(cudaMemcpyAsync(fitnesses + r0, s.fitness[b], nr * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    EVOGP_CUDA(cudaStreamSynchronize(s.stream[0]));
    EVOGP_CUDA(cudaStreamSynchronize(s.stream[1]));
    return EVOGP_OK;
}
