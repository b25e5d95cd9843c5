// splice.cu — subtree crossover and subtree mutation over the packed arrays.
//
// Replaces crossover() / mutate() (src/evogp/cuda/mutation.cu:186-219, :312-347),
// their kernels (:118-184, :224-309) and the serial splice routine _gpTreeReplace
// (:5-115).  The reference runs one THREAD per child, staging three 1024-entry
// arrays in local memory and copying them element by element.  Here one WARP builds
// one child row and nothing is staged: every output slot j knows in O(1) where it
// comes from,
//
//      j <  pos                 recipient[j]         (+ size fix-up, below)
//      j <  pos + dsize         donor[dpos + j - pos]
//      j <  newlen              recipient[j - diff]
//      else                     0                    (tail is zero-filled)
//
// and the reference's root->pos walk that adds `diff` to every ancestor's
// subtree_size (:38-88) collapses to the prefix-order identity
//      j is an ancestor of pos  <=>  j < pos < j + size[j].
// Lanes take adjacent slot pairs, so the i16 arrays are written as 32-bit words and
// node_value as 64-bit words: full-row, fully coalesced stores.
#include "common.cuh"

namespace evogp {

struct SpliceArgs {
    const float *value;      // recipient rows  [P_src][L]
    const int16_t *type;
    const int16_t *size;
    const float *dvalue;     // donor rows (crossover: same arrays as recipient)
    const int16_t *dtype;
    const int16_t *dsize;
    const int *left_idx;     // crossover: recipient row per child;   mutation: nullptr (row n)
    const int *right_idx;    // crossover: donor row per child;       mutation: nullptr (row n)
    const int *left_node;    // splice position in the recipient
    const int *right_node;   // subtree root in the donor;            mutation: nullptr (0)
    float *ovalue;
    int16_t *otype;
    int16_t *osize;
    int P_src, P_new, L;
};

template <bool CROSSOVER>
__global__ void __launch_bounds__(256, 6) splice_kernel(SpliceArgs g) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= g.P_new) return;
    const int L = g.L;

    // ---- per-child header (warp-uniform).  The kernel is bound by the latency of its dependent loads (ncu: issue slots
    //      42 % busy, DRAM 21 %), so the header is two batches of independent loads: indices, then the three sizes it
    //      needs (positions clamped so that the loads can leave before the range checks are known) ----
    int lrow = n, rrow = n, pos, dpos = 0;
    if (CROSSOVER) {
        lrow = __ldg(g.left_idx + n);
        rrow = __ldg(g.right_idx + n);
        dpos = __ldg(g.right_node + n);
    }
    pos = __ldg(g.left_node + n);
    if (CROSSOVER && (lrow < 0 || lrow >= g.P_src)) lrow = 0;   // reference: undefined behaviour
    const bool rrow_ok = !CROSSOVER || (rrow >= 0 && rrow < g.P_src);
    const float *lv = g.value + (size_t)lrow * L;
    const int16_t *lt = g.type + (size_t)lrow * L;
    const int16_t *ls = g.size + (size_t)lrow * L;
    const float *rv = g.dvalue + (size_t)(rrow_ok ? rrow : 0) * L;
    const int16_t *rt = g.dtype + (size_t)(rrow_ok ? rrow : 0) * L;
    const int16_t *rs = g.dsize + (size_t)(rrow_ok ? rrow : 0) * L;
    const bool pos_in_row = pos >= 0 && pos < L, dpos_in_row = dpos >= 0 && dpos < L;
    const int left_size = __ldg(ls);
    const int lsub_raw = __ldg(ls + (pos_in_row ? pos : 0));
    const int dsub_raw = __ldg(rs + (dpos_in_row ? dpos : 0));
    // mutation.cu:256 (donor row range) and :150 (position range).  The reference does not
    // range-check crossover positions (undefined behaviour there); here they fall back to a copy.
    bool ok = pos >= 0 && pos < left_size && dpos_in_row && rrow_ok;
    int lsub = 0, dsub = 0, diff = 0;
    if (ok) {
        lsub = lsub_raw;
        dsub = dsub_raw;
        diff = dsub - lsub;
        ok = dsub >= 1 && left_size + diff <= L;                     // mutation.cu:163, :279
    }
    if (!ok) {   // reference falls back to a plain copy of the recipient
        pos = left_size;
        dsub = 0;
        diff = 0;
    }
    const int newlen = left_size + diff;
    const int dend = pos + dsub;

    float *ov = g.ovalue + (size_t)n * L;
    int16_t *ot = g.otype + (size_t)n * L;
    int16_t *os = g.osize + (size_t)n * L;

    auto gather = [&](int j, float &v, int &t, int &s) {
        if (j >= newlen) { v = 0.0f; t = 0; s = 0; return; }
        if (j < pos) {
            v = __ldg(lv + j); t = __ldg(lt + j); s = __ldg(ls + j);
            if (j + s > pos) s += diff;          // ancestor of the splice point
        } else if (j < dend) {
            const int q = dpos + j - pos;
            v = __ldg(rv + q); t = __ldg(rt + q); s = __ldg(rs + q);
        } else {
            const int q = j - diff;
            v = __ldg(lv + q); t = __ldg(lt + q); s = __ldg(ls + q);
        }
    };

    if ((L & 1) == 0) {
        for (int j = lane * 2; j < L; j += 64) {
            float v0, v1;
            int t0, t1, s0, s1;
            gather(j, v0, t0, s0);
            gather(j + 1, v1, t1, s1);
            *reinterpret_cast<float2 *>(ov + j) = make_float2(v0, v1);
            *reinterpret_cast<uint32_t *>(ot + j) = (uint32_t)(uint16_t)t0 | ((uint32_t)(uint16_t)t1 << 16);
            *reinterpret_cast<uint32_t *>(os + j) = (uint32_t)(uint16_t)s0 | ((uint32_t)(uint16_t)s1 << 16);
        }
    } else {
        for (int j = lane; j < L; j += 32) {
            float v;
            int t, s;
            gather(j, v, t, s);
            ov[j] = v; ot[j] = (int16_t)t; os[j] = (int16_t)s;
        }
    }
}

template <bool CROSSOVER>
static int launch_splice(const SpliceArgs &a, cudaStream_t st) {
    const int warps = 8;
    const int grid = (a.P_new + warps - 1) / warps;
    splice_kernel<CROSSOVER><<<grid, warps * 32, 0, st>>>(a);
    count_launch();
    return check_launch(CROSSOVER ? "crossover" : "mutate");
}

}  // namespace evogp

using namespace evogp;

extern "C" int evogp_crossover(int pop_size_ori, int pop_size_new, int gpLen, const float *value_ori,
                               const int16_t *type_ori, const int16_t *subtree_size_ori, const int *left_idx,
                               const int *right_idx, const int *left_node_idx, const int *right_node_idx,
                               float *value_res, int16_t *type_res, int16_t *subtree_size_res, void *stream) {
    EVOGP_REQUIRE(pop_size_ori > 0, "pop_size_ori must be larger than 0, got %d", pop_size_ori);
    EVOGP_REQUIRE(pop_size_new > 0, "pop_size_new must be larger than 0, got %d", pop_size_new);
    EVOGP_REQUIRE(gpLen > 0 && gpLen <= kMaxStack, "gp_len must be in (0, %d], got %d", kMaxStack, gpLen);
    int rc = ensure_device_ok();
    if (rc) return rc;
    SpliceArgs a;
    a.value = value_ori; a.type = type_ori; a.size = subtree_size_ori;
    a.dvalue = value_ori; a.dtype = type_ori; a.dsize = subtree_size_ori;
    a.left_idx = left_idx; a.right_idx = right_idx; a.left_node = left_node_idx; a.right_node = right_node_idx;
    a.ovalue = value_res; a.otype = type_res; a.osize = subtree_size_res;
    a.P_src = pop_size_ori; a.P_new = pop_size_new; a.L = gpLen;
    return launch_splice<true>(a, static_cast<cudaStream_t>(stream));
}

extern "C" int evogp_mutate(int popSize, int gpLen, const float *value_ori, const int16_t *type_ori,
                            const int16_t *subtree_size_ori, const int *mutateIndices, const float *value_new,
                            const int16_t *type_new, const int16_t *subtree_size_new, float *value_res,
                            int16_t *type_res, int16_t *subtree_size_res, void *stream) {
    EVOGP_REQUIRE(popSize > 0, "pop_size must be larger than 0, got %d", popSize);
    EVOGP_REQUIRE(gpLen > 0 && gpLen <= kMaxStack, "gp_len must be in (0, %d], got %d", kMaxStack, gpLen);
    int rc = ensure_device_ok();
    if (rc) return rc;
    SpliceArgs a;
    a.value = value_ori; a.type = type_ori; a.size = subtree_size_ori;
    a.dvalue = value_new; a.dtype = type_new; a.dsize = subtree_size_new;
    a.left_idx = nullptr; a.right_idx = nullptr; a.left_node = mutateIndices; a.right_node = nullptr;
    a.ovalue = value_res; a.otype = type_res; a.osize = subtree_size_res;
    a.P_src = popSize; a.P_new = popSize; a.L = gpLen;
    return launch_splice<false>(a, static_cast<cudaStream_t>(stream));
}
