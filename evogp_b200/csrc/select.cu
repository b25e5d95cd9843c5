// select.cu — operator variants on the same boundary.
//
// extract_subtree: the native form of vmap_subtree / subtensor (src/evogp/algorithm/mutation/mutation_utils.py:6-48),
//   which Hoist / Insert / Delete mutation build from gather + where over three [P, L] tensors.  Row n of the result is
//   the subtree of tree n rooted at pos[n], moved to the front, tail zero-filled: one warp per row, one pass.
// tournament_select: TournamentSelection (src/evogp/algorithm/selection/tournament.py:59-133).  The reference draws the
//   contenders with torch.multinomial over a [k_times, P] matrix of ones under vmap (:73-79, :116-120) and picks each
//   winner with argsort under vmap (:90-103).  Here one thread runs one tournament: contenders come from a counter-based
//   generator (Philox4x32-10 keyed by `keys`, the tournament index and the draw index), "without replacement" from a
//   keyed bijection of [0, P) (4-round Feistel network + cycle walking, one permutation per round of P / t_size
//   tournaments) instead of a materialised permutation, and the winner is the nth best contender with
//   nth = floor(log u / log(1 - best_p)) (:97-101).  Same distribution, not torch's random stream; bit-exact against the
//   oracle's restatement (oracle/evogp_oracle.c: oracle_tournament).
#include "gen_tree.cuh"

namespace evogp {

struct ExtractArgs {
    const float *value;
    const int16_t *type;
    const int16_t *size;
    const int *pos;
    float *ovalue;
    int16_t *otype;
    int16_t *osize;
    int P, L;
};

__global__ void __launch_bounds__(256) extract_subtree_kernel(ExtractArgs g) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= g.P) return;
    const int L = g.L;
    const size_t row = (size_t)n * L;
    int p = __ldg(g.pos + n);
    // mutation_utils.py:21-24: indices are clamped to the row, entries at or beyond start + length become 0
    const bool ok = p >= 0 && p < L;
    const int len = ok ? (int)__ldg(g.size + row + p) : 0;
    for (int j = lane; j < L; j += 32) {
        const bool in = ok && j < len && p + j < L;
        g.ovalue[row + j] = in ? __ldg(g.value + row + p + j) : 0.0f;
        g.otype[row + j] = in ? __ldg(g.type + row + p + j) : (int16_t)0;
        g.osize[row + j] = in ? __ldg(g.size + row + p + j) : (int16_t)0;
    }
}

// ---- keyed bijection of [0, n): 4-round Feistel network on 2 * half bits, cycle-walked into range ----
constexpr uint32_t kTournamentStream = 0x20000u, kPermStream = 0x30000u;

__host__ __device__ __forceinline__ uint32_t feistel_perm(uint32_t x, uint32_t n, int half_bits, const uint32_t rk[4]) {
    const uint32_t mask = (1u << half_bits) - 1u;
    do {
        uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t f = (r ^ rk[i]) * 0x9E3779B1u;      // round function: keyed multiplicative hash of the right half
            f ^= f >> 15;
            f *= 0x85EBCA77u;
            f ^= f >> 13;
            const uint32_t nl = r;
            r = (l ^ f) & mask;
            l = nl;
        }
        x = (l << half_bits) | r;
    } while (x >= n);
    return x;
}

struct TournamentArgs {
    const float *fitness;
    const unsigned *keys;
    int *winners;
    int P, t_size, count, replace;
    float best_p;
};

__global__ void __launch_bounds__(256) tournament_kernel(TournamentArgs g) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.count) return;
    const uint32_t k0 = g.keys[0], k1 = g.keys[1];
    const uint32_t P = (uint32_t)g.P;
    const int T = g.t_size;
    // nth best wins with probability best_p * (1 - best_p)^nth, i.e. nth = floor(log u / log(1 - best_p)), and 0 when
    // that overshoots the tournament (tournament.py:97-101).  Computed with a running product instead of logarithms so
    // that the CPU oracle reproduces it bit for bit:  (1 - p)^(nth + 1) < u <= (1 - p)^nth
    uint32_t d[4];
    philox4x32_10((uint32_t)j, kTournamentStream, k0, k1, d);
    int nth = 0;
    if (g.best_p < 1.0f) {
        const float u = __uint2float_rn(d[0]) * 2.3283064365386963e-10f, q = __fsub_rn(1.0f, g.best_p);
        float thr = q;
        while (nth < T && u <= thr) {
            ++nth;
            thr = __fmul_rn(thr, q);
        }
        if (nth >= T) nth = 0;
    }
    // without replacement: tournaments are consecutive slices of a fresh pseudo-random permutation per round
    const int per_round = g.P / T;                      // tournaments one permutation serves (tournament.py:113)
    const int round = j / per_round, slot = j - round * per_round;
    int half_bits = 1;
    while ((1ull << (2 * half_bits)) < (unsigned long long)P) ++half_bits;
    uint32_t rk[4];
    philox4x32_10((uint32_t)round, kPermStream, k0, k1, rk);
    auto contender = [&](int i) -> uint32_t {
        if (g.replace) {
            uint32_t w[4];
            philox4x32_10((uint32_t)j, kTournamentStream + 1u + (uint32_t)(i >> 2), k0, k1, w);
            const uint32_t x = (i & 3) == 0 ? w[0] : ((i & 3) == 1 ? w[1] : ((i & 3) == 2 ? w[2] : w[3]));
            return x % P;
        }
        return feistel_perm((uint32_t)(slot * T + i), P, half_bits, rk);
    };
    auto key = [&](uint32_t c) -> float {
        const float f = __ldg(g.fitness + c);
        return f == f ? f : -__int_as_float(0x7f800000);   // NaN ranks last (pipeline/standard.py:43 maps it to -inf)
    };
    // selection by counting: contender i wins iff exactly `nth` contenders beat it (ties: the earlier draw is better)
    constexpr int CACHE = 32;
    uint32_t cs[CACHE];
    float fs[CACHE];
    for (int i = 0; i < T && i < CACHE; ++i) {
        cs[i] = contender(i);
        fs[i] = key(cs[i]);
    }
    int win = 0;
    for (int i = 0; i < T; ++i) {
        const uint32_t ci = i < CACHE ? cs[i] : contender(i);
        const float fi = i < CACHE ? fs[i] : key(ci);
        int better = 0;
        for (int m = 0; m < T; ++m) {
            if (m == i) continue;
            const float fm = m < CACHE ? fs[m] : key(contender(m));
            better += (fm > fi || (fm == fi && m < i)) ? 1 : 0;
        }
        if (better == nth) win = (int)ci;
    }
    g.winners[j] = win;
}

// ---- multi-GPU: this rank's fitness slice -> every rank's full-population buffer (peer-mapped memory, NVLink) ----
__global__ void __launch_bounds__(256) push_fitness_kernel(const float *__restrict__ local, unsigned count, float *const *peers,
                                                           int world, unsigned row_offset) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (int r = 0; r < world; ++r) {
        float *dst = peers[r] + row_offset;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[i] = local[i];   // 128 B per warp store
    }
}

}  // namespace evogp

using namespace evogp;

extern "C" int evogp_push_fitness(const float *local_fitness, unsigned count, float *const *peer_fitnesses, unsigned world,
                                  unsigned row_offset, void *stream) {
    EVOGP_REQUIRE(local_fitness != nullptr && peer_fitnesses != nullptr, "pointers must not be NULL");
    EVOGP_REQUIRE(world >= 1 && world <= 64, "world must be in [1, 64], got %u", world);
    if (count == 0) return EVOGP_OK;
    int rc = ensure_device_ok();
    if (rc) return rc;
    unsigned grid = (count + 255) / 256;
    if (grid > 296) grid = 296;
    push_fitness_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(local_fitness, count, peer_fitnesses, (int)world, row_offset);
    count_launch();
    return check_launch("push_fitness");
}

extern "C" int evogp_extract_subtree(int popSize, int gpLen, const float *value, const int16_t *type,
                                     const int16_t *subtree_size, const int *positions, float *value_res,
                                     int16_t *type_res, int16_t *subtree_size_res, void *stream) {
    EVOGP_REQUIRE(popSize > 0, "pop_size must be larger than 0, got %d", popSize);
    EVOGP_REQUIRE(gpLen > 0 && gpLen <= kMaxStack, "gp_len must be in (0, %d], got %d", kMaxStack, gpLen);
    int rc = ensure_device_ok();
    if (rc) return rc;
    ExtractArgs a;
    a.value = value; a.type = type; a.size = subtree_size; a.pos = positions;
    a.ovalue = value_res; a.otype = type_res; a.osize = subtree_size_res;
    a.P = popSize; a.L = gpLen;
    const int warps = 8;
    extract_subtree_kernel<<<(popSize + warps - 1) / warps, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
    count_launch();
    return check_launch("extract_subtree");
}

extern "C" int evogp_tournament_select(int popSize, const float *fitness, int tournamentSize, float bestProbability,
                                       int replace, int winnerCnt, const unsigned *keys, int *winners, void *stream) {
    EVOGP_REQUIRE(popSize > 0, "pop_size must be larger than 0, got %d", popSize);
    EVOGP_REQUIRE(tournamentSize > 0 && tournamentSize <= popSize, "tournament_size must be in [1, pop_size], got %d", tournamentSize);
    EVOGP_REQUIRE(bestProbability > 0.0f && bestProbability <= 1.0f, "best_probability must be in (0, 1], got %f", bestProbability);
    EVOGP_REQUIRE(winnerCnt > 0, "winner_cnt must be larger than 0, got %d", winnerCnt);
    int rc = ensure_device_ok();
    if (rc) return rc;
    TournamentArgs a;
    a.fitness = fitness; a.keys = keys; a.winners = winners;
    a.P = popSize; a.t_size = tournamentSize; a.count = winnerCnt; a.replace = replace ? 1 : 0; a.best_p = bestProbability;
    tournament_kernel<<<(winnerCnt + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    count_launch();
    return check_launch("tournament_select");
}
