// gen_tree.cuh — per-tree random generation shared by generate.cu and nextgen.cu.
// Bit-identical to the reference's treeGPGenerate draw sequence (src/evogp/cuda/generate.cu:16-173):
// FNV-1a seed (kernel.h:157-180), thrust taus88, same draw order, downward roulette scan.
#pragma once
#include "common.cuh"

namespace evogp {

// kernel.h:160-172: low 32 bits of 64-bit FNV-1a over the bytes of {n, k1, k2}.  Only the low word is used, and the low word
// of (h ^ byte) * prime depends on the low words alone: the hash runs in 32-bit arithmetic (offset basis
// 0xCBF29CE484222325, prime 0x100000001B3 -> their low halves).
__device__ __forceinline__ uint32_t tree_seed(uint32_t n, uint32_t k1, uint32_t k2) {
    uint32_t h = 0x84222325u;
    const uint32_t a[3] = {n, k1, k2};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            h ^= (a[i] >> (8 * b)) & 0xFFu;
            h *= 0x000001B3u;
        }
    return h;
}

// thrust::random::taus88 (kernel.h:20): three LFSRs, all seeded with the same word
struct Taus88 {
    uint32_t z1, z2, z3;
    __device__ __forceinline__ explicit Taus88(uint32_t s) : z1(s), z2(s), z3(s) {}
    __device__ __forceinline__ uint32_t next() {
        uint32_t b;
        b = ((z1 << 13) ^ z1) >> 19;
        z1 = ((z1 & 0xFFFFFFFEu) << 12) ^ b;
        b = ((z2 << 2) ^ z2) >> 25;
        z2 = ((z2 & 0xFFFFFFF8u) << 4) ^ b;
        b = ((z3 << 3) ^ z3) >> 11;
        z3 = ((z3 & 0xFFFFFFF0u) << 17) ^ b;
        return z1 ^ z2 ^ z3;
    }
    // thrust::uniform_real_distribution<float>(0,1): float(u32) / 2^32 (exact scaling; can return 1.0f)
    __device__ __forceinline__ float uniform() { return __uint2float_rn(next()) * 2.3283064365386963e-10f; }
};


// Philox4x32-10 (Salmon et al., SC'11): counter (c0, c1, 0, 0), key (k0, k1)
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t c2 = 0, c3 = 0;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Counter-based per-tree stream for evogp_generate's Philox mode (BASELINE.json north_star: "cuRAND/Philox per-thread
// state"; the reference's RNG site is generate.cu:40-41): draw j of tree n is word j % 4 of
// philox4x32_10(counter = (n, kPhiloxGenStream + j / 4), key = keys).  No state in memory, any tree reproducible alone.
constexpr uint32_t kPhiloxGenStream = 0x10000u;
struct PhiloxStream {
    uint32_t n, k0, k1, blk, buf[4];
    int have;
    __device__ __forceinline__ PhiloxStream(uint32_t n_, uint32_t k0_, uint32_t k1_) : n(n_), k0(k0_), k1(k1_), blk(0), have(0) {}
    __device__ __forceinline__ uint32_t next() {
        if (have == 0) {
            philox4x32_10(n, kPhiloxGenStream + blk, k0, k1, buf);
            ++blk;
            have = 4;
        }
        const int i = 4 - have;
        --have;
        return i == 0 ? buf[0] : (i == 1 ? buf[1] : (i == 2 ? buf[2] : buf[3]));
    }
    __device__ __forceinline__ float uniform() { return __uint2float_rn(next()) * 2.3283064365386963e-10f; }
};

struct GrowParams {
    const float *leaf;     // [10]  depth -> leaf probability   (shared or global memory)
    const float *roul;     // [29]  cumulative function roulette
    const float *consts;   // [S]   global memory
    unsigned L, V, O, S;
    float outProb, constProb;
};

// Grows one tree into val[] (value bits) and ts[] (type | size << 16); returns the node count.
// The frame stack is a register: frames have strictly increasing depth, so "children still owed at depth d"
// is a 4-bit field of one 64-bit word.  Subtree sizes need no stack: scanning the prefix backwards,
// size[i] = 1 + size[c1] + size[c2] + ... with c1 = i + 1, c2 = c1 + size[c1].
template <bool MULTI, class Rng = Taus88>
__device__ inline int grow_tree(Rng &rng, const GrowParams &g, uint32_t *val, uint32_t *ts) {
    uint64_t owed = 1;   // root frame {1, 0}
    int d = 0, cnt = 0;
    while (d >= 0 && cnt < (int)g.L) {
        owed -= 1ull << (4 * d);                                   // cd.childs-- (generate.cu:61)
        const float leafp = d < kMaxFullDepth ? g.leaf[d] : 2.0f;  // reference indexes out of bounds at d >= 10
        uint32_t vbits;
        int type, arity = 0;
        if (rng.uniform() >= leafp) {                              // function node (:71)
            const float r = rng.uniform();
            int k = 0;
            for (int i = F_END - 1; i >= 0; --i)                   // downward roulette scan (:74-84)
                if (r >= g.roul[i]) { k = i + 1; break; }
            type = k <= F_IF ? NT_TFUNC : (k <= F_GE ? NT_BFUNC : NT_UFUNC);
            arity = type - 1;
            vbits = __float_as_uint((float)k);
            if (MULTI) {
                if (rng.uniform() <= g.outProb) {                  // output node (:88-96)
                    const uint32_t oi = rng.next() % g.O;
                    vbits = ((uint32_t)k & 0xFFFFu) | (oi << 16);  // kernel.h:105-113
                    type += NT_OUT;
                }
            }
        } else if (rng.uniform() <= g.constProb) {                 // constant leaf (:109-114)
            vbits = __float_as_uint(__ldg(g.consts + rng.next() % g.S));
            type = NT_CONST;
        } else {                                                   // variable leaf (:116-120)
            vbits = __float_as_uint((float)(rng.next() % g.V));
            type = NT_VAR;
        }
        val[cnt] = vbits;
        ts[cnt] = (uint32_t)type & 0xFFFFu;
        ++cnt;
        if (arity > 0 && d + 1 < 16) {
            ++d;
            owed |= (uint64_t)arity << (4 * d);
        } else {
            while (d >= 0 && ((owed >> (4 * d)) & 0xF) == 0) --d;
        }
    }
    for (int i = cnt - 1; i >= 0; --i) {                           // subtree sizes, leaves -> root (:130-158)
        const int t = ts[i] & NT_MASK;
        const int ar = t <= NT_CONST ? 0 : t - 1;
        int sz = 1, c = i + 1;
        for (int k = 0; k < ar; ++k) {
            const int cs = c < cnt ? (int)(ts[c] >> 16) : 0;
            sz += cs;
            c += cs;
        }
        ts[i] |= (uint32_t)sz << 16;
    }
    return cnt;
}


// ---------------------------------------------------------------------------
// Packed, branch-free growth (single-output trees, taus88 draws) - shared by generate_fast_kernel (generate.cu) and the
// donor phase of nextgen_kernel (nextgen.cu).  Same draws, same trees as grow_tree<false, Taus88>.
// ---------------------------------------------------------------------------
struct Taus88State {
    uint32_t z1, z2, z3;
};
__device__ __forceinline__ uint32_t taus88_step(Taus88State &s) {
    uint32_t b;
    b = ((s.z1 << 13) ^ s.z1) >> 19;
    s.z1 = ((s.z1 & 0xFFFFFFFEu) << 12) ^ b;
    b = ((s.z2 << 2) ^ s.z2) >> 25;
    s.z2 = ((s.z2 & 0xFFFFFFF8u) << 4) ^ b;
    b = ((s.z3 << 3) ^ s.z3) >> 11;
    s.z3 = ((s.z3 & 0xFFFFFFF0u) << 17) ^ b;
    return s.z1 ^ s.z2 ^ s.z3;
}
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return __uint2float_rn(x) * 2.3283064365386963e-10f; }   // float(u32) / 2^32

// x % d for a divisor known before the loop: Lemire's fastmod, M = floor(2^64 / d) + 1 computed on the host
__device__ __forceinline__ uint32_t fastmod_u32(uint32_t x, uint64_t M, uint32_t d) {
    return (uint32_t)__umul64hi(M * x, (uint64_t)d);
}


// Node word: [2:0] type, [15:4] subtree size, [31:16] variable / constant-sample index (leaves); functions keep their
// id in [20:16] and, while their frame is open, the index of the enclosing function node in [31:21].
// s_leaf: 16 floats (depth -> leaf probability, 2.0 beyond MAX_FULL_DEPTH); s_roul: 32 floats (cumulative roulette padded
// with +inf); mono: the roulette is non-decreasing (binary search allowed).  Returns the tree length (0 when !active).
// One tree being grown: the registers of the loop below (generate_fast_kernel and nextgen grow a tree to the end with
// grow_tree_packed; generate_balanced_kernel steps 32 of them side by side and re-arms a lane when its tree is done).
struct PackedGrowth {
    int cnt, d;
    uint32_t owed;                     // children still owed per depth, 2 bits each (arity <= 3); root frame {1, 0}
    uint32_t cur;                      // the function node whose children are being generated (depth >= 1)
    Taus88State st;
    __device__ __forceinline__ void start(uint32_t seed, bool active) {
        cnt = 0; d = active ? 0 : -1; owed = 1; cur = 0;
        st.z1 = st.z2 = st.z3 = seed;
    }
    __device__ __forceinline__ bool growing(int L) const { return d >= 0 && cnt < L; }
    // one node (generate.cu:58-128 of the reference, one iteration of its loop)
    __device__ __forceinline__ void step(const float *s_leaf, const float *s_roul, bool mono, uint32_t V, uint32_t S, uint64_t MV,
                                         uint64_t MS, float constProb, uint32_t *row) {
        owed -= 1u << (2 * d);                                     // cd.childs-- (generate.cu:61)
        const float leafp = s_leaf[d];
        // draws: u (leaf test); then r (roulette) or u (constant test) - the same word; then, for a leaf only, a raw word.
        // All three are made; a function commits the state after two.
        Taus88State s1 = st;
        const uint32_t o1 = taus88_step(s1);
        Taus88State s2 = s1;
        const uint32_t o2 = taus88_step(s2);
        Taus88State s3 = s2;
        const uint32_t o3 = taus88_step(s3);
        const bool is_func = u32_to_unit(o1) >= leafp;             // :71
        const float r = u32_to_unit(o2);
        int k = 0;                                                 // number of roulette entries <= r (:74-84)
        if (mono) {
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (s_roul[k + step - 1] <= r) k += step;
        } else {
            for (int i = F_END - 1; i >= 0; --i)
                if (r >= s_roul[i]) { k = i + 1; break; }
        }
        const uint32_t ftype = k <= F_IF ? NT_TFUNC : (k <= F_GE ? NT_BFUNC : NT_UFUNC);
        const bool is_const = r <= constProb;                      // :109
        const uint32_t idx = is_const ? fastmod_u32(o3, MS, S) : fastmod_u32(o3, MV, V);
        st.z1 = is_func ? s2.z1 : s3.z1;
        st.z2 = is_func ? s2.z2 : s3.z2;
        st.z3 = is_func ? s2.z3 : s3.z3;
        if (is_func) {                                             // open the frame of its children
            row[cnt] = ((uint32_t)k << 16) | (cur << 21) | ftype;
            cur = (uint32_t)cnt;
            ++d;
            owed |= (ftype - 1u) << (2 * d);
            ++cnt;
        } else {
            row[cnt] = (idx << 16) | (1u << 4) | (is_const ? (uint32_t)NT_CONST : (uint32_t)NT_VAR);
            ++cnt;
            while (d >= 0 && ((owed >> (2 * d)) & 3u) == 0u) {      // frames whose children are all there: their node is complete
                if (d > 0) {
                    const uint32_t w = row[cur];
                    row[cur] = (w & 0x001FFFFFu) | ((uint32_t)(cnt - (int)cur) << 4);
                    cur = w >> 21;
                }
                --d;
            }
        }
    }
    // row full before the tree closed (a descriptor check_tree_length would have refused): close what is open.
    // Returns the tree length (0 for a lane that never had a tree).
    __device__ __forceinline__ int finish(uint32_t *row) {
        for (; d > 0; --d) {
            const uint32_t w = row[cur];
            row[cur] = (w & 0x001FFFFFu) | ((uint32_t)(cnt - (int)cur) << 4);
            cur = w >> 21;
        }
        return cnt > 0 ? (int)((row[0] >> 4) & 0xFFF) : 0;
    }
};

__device__ __forceinline__ int grow_tree_packed(uint32_t seed, bool active, const float *s_leaf, const float *s_roul, bool mono,
                                                uint32_t V, uint32_t S, uint64_t MV, uint64_t MS, float constProb, int L,
                                                uint32_t *row) {
    PackedGrowth t;
    t.start(seed, active);
    while (t.growing(L)) t.step(s_leaf, s_roul, mono, V, S, MV, MS, constProb, row);
    return t.finish(row);
}

// packed node word -> value bits, node type, subtree size.  Branch-free (the constant table is read at index 0 for the
// other node types); a zero word - the padding behind a tree - decodes to three zeros.
__device__ __forceinline__ void decode_packed_node(uint32_t w, const float *consts, uint32_t &v, uint32_t &t, uint32_t &sz) {
    t = w & 7u;
    sz = (w >> 4) & 0xFFFu;
    const uint32_t code = w >> 16;
    const float cv = __ldg(consts + (t == NT_CONST ? code : 0u));
    const float fv = (float)(t == NT_VAR ? code : (code & 31u));
    v = __float_as_uint(t == NT_CONST ? cv : fv);
}

// One packed row (len valid words at src, in shared memory) -> one zero-filled row of the three output arrays, by a warp.
__device__ __forceinline__ void write_packed_row(const uint32_t *src, int len, int lane, int L, const float *consts, float *ov,
                                                 int16_t *ot, int16_t *os) {
    if ((L & 1) == 0) {
        for (int j = lane * 2; j < L; j += 64) {
            const uint32_t w0 = j < len ? src[j] : 0u, w1 = j + 1 < len ? src[j + 1] : 0u;
            uint32_t v0, t0, z0, v1, t1, z1;
            decode_packed_node(w0, consts, v0, t0, z0);
            decode_packed_node(w1, consts, v1, t1, z1);
            *reinterpret_cast<uint2 *>(ov + j) = make_uint2(v0, v1);
            *reinterpret_cast<uint32_t *>(ot + j) = t0 | (t1 << 16);
            *reinterpret_cast<uint32_t *>(os + j) = z0 | (z1 << 16);
        }
    } else {
        for (int j = lane; j < L; j += 32) {
            uint32_t v, t, z;
            decode_packed_node(j < len ? src[j] : 0u, consts, v, t, z);
            ov[j] = __uint_as_float(v);
            ot[j] = (int16_t)t;
            os[j] = (int16_t)z;
        }
    }
}

}  // namespace evogp
