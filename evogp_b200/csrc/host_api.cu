// host_api.cu — host-buffer entry point (evogp_SR_fitness_host): the whole
// "forest on the host -> fitness on the host" trip a non-torch caller makes.
// The forest is cut into row chunks that alternate between two streams, so the
// H2D copy of chunk c+1 overlaps the lowering + replay of chunk c and the D2H of
// its fitness slice.  Only what the evaluator reads crosses PCIe: node_value and
// node_type rows plus ONE length per tree (column 0 of subtree_size, gathered on
// the host into a pinned staging array) — 6 B per node slot instead of 8.
// Staging buffers are cached across calls.
#include <cstdlib>
#include <mutex>
#include "common.cuh"

int evogp_sr_fitness_compact_len(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen, unsigned outLen,
                                 int useMSE, const float *value, const int16_t *type, const int16_t *lengths,
                                 const float *variables, const float *labels, float *fitnesses, void *workspace,
                                 size_t workspace_bytes, void *stream);

namespace evogp {
namespace {

struct Staging {
    int device = -1;
    size_t chunk_rows = 0, L = 0;
    float *value[2] = {nullptr, nullptr};
    int16_t *type[2] = {nullptr, nullptr};
    int16_t *len[2] = {nullptr, nullptr};   // device: one length per tree of the chunk
    int16_t *host_len = nullptr;            // pinned: lengths of the whole population (column 0 of subtree_size)
    size_t host_len_cap = 0;
    void *ws[2] = {nullptr, nullptr};
    size_t ws_bytes = 0;
    float *fitness[2] = {nullptr, nullptr};
    float *X = nullptr, *labels = nullptr;
    size_t x_cap = 0, lab_cap = 0;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t data_ready = nullptr;
};
Staging g_st;
std::mutex g_mu;

void release_locked() {
    if (g_st.device < 0) return;
    cudaSetDevice(g_st.device);
    for (int i = 0; i < 2; ++i) {
        cudaFree(g_st.value[i]); cudaFree(g_st.type[i]); cudaFree(g_st.len[i]);
        cudaFree(g_st.ws[i]); cudaFree(g_st.fitness[i]);
        if (g_st.stream[i]) cudaStreamDestroy(g_st.stream[i]);
    }
    cudaFree(g_st.X); cudaFree(g_st.labels);
    if (g_st.host_len) cudaFreeHost(g_st.host_len);
    if (g_st.data_ready) cudaEventDestroy(g_st.data_ready);
    g_st = Staging();
}

int prepare(int device, size_t rows, size_t L, size_t xbytes, size_t lbytes) {
    if (g_st.device != device || g_st.chunk_rows < rows || g_st.L != L) {
        release_locked();
        EVOGP_CUDA(cudaSetDevice(device));
        g_st.device = device;
        g_st.chunk_rows = rows;
        g_st.L = L;
        g_st.ws_bytes = evogp_eval_workspace_bytes((unsigned)rows, (unsigned)L);
        for (int i = 0; i < 2; ++i) {
            EVOGP_CUDA(cudaMalloc(&g_st.value[i], rows * L * sizeof(float)));
            EVOGP_CUDA(cudaMalloc(&g_st.type[i], rows * L * sizeof(int16_t)));
            EVOGP_CUDA(cudaMalloc(&g_st.len[i], rows * sizeof(int16_t)));
            EVOGP_CUDA(cudaMalloc(&g_st.ws[i], g_st.ws_bytes));
            EVOGP_CUDA(cudaMalloc(&g_st.fitness[i], rows * sizeof(float)));
            EVOGP_CUDA(cudaStreamCreateWithFlags(&g_st.stream[i], cudaStreamNonBlocking));
        }
        EVOGP_CUDA(cudaEventCreateWithFlags(&g_st.data_ready, cudaEventDisableTiming));
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
    if (s.host_len_cap < popSize) {
        if (s.host_len) cudaFreeHost(s.host_len);
        EVOGP_CUDA(cudaHostAlloc(&s.host_len, (size_t)popSize * sizeof(int16_t), cudaHostAllocDefault));
        s.host_len_cap = popSize;
    }
    EVOGP_CUDA(cudaMemcpyAsync(s.X, variables, xbytes, cudaMemcpyHostToDevice, s.stream[0]));
    EVOGP_CUDA(cudaMemcpyAsync(s.labels, labels, lbytes, cudaMemcpyHostToDevice, s.stream[0]));
    EVOGP_CUDA(cudaEventRecord(s.data_ready, s.stream[0]));
    EVOGP_CUDA(cudaStreamWaitEvent(s.stream[1], s.data_ready, 0));
    int c = 0;
    for (size_t r0 = 0; r0 < popSize; r0 += rows, ++c) {
        const int b = c & 1;
        const size_t nr = (popSize - r0 < rows) ? popSize - r0 : rows;
        cudaStream_t st = s.stream[b];
        EVOGP_CUDA(cudaMemcpyAsync(s.value[b], value + r0 * L, nr * L * sizeof(float), cudaMemcpyHostToDevice, st));
        EVOGP_CUDA(cudaMemcpyAsync(s.type[b], type + r0 * L, nr * L * sizeof(int16_t), cudaMemcpyHostToDevice, st));
        for (size_t r = 0; r < nr; ++r) s.host_len[r0 + r] = subtree_size[(r0 + r) * L];   // the only column the evaluator reads
        EVOGP_CUDA(cudaMemcpyAsync(s.len[b], s.host_len + r0, nr * sizeof(int16_t), cudaMemcpyHostToDevice, st));
        rc = evogp_sr_fitness_compact_len((unsigned)nr, dataPoints, gpLen, varLen, outLen, useMSE, s.value[b], s.type[b],
                                          s.len[b], s.X, s.labels, s.fitness[b], s.ws[b], s.ws_bytes, st);
        if (rc) return rc;
        EVOGP_CUDA(cudaMemcpyAsync(fitnesses + r0, s.fitness[b], nr * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    EVOGP_CUDA(cudaStreamSynchronize(s.stream[0]));
    EVOGP_CUDA(cudaStreamSynchronize(s.stream[1]));
    return EVOGP_OK;
}
