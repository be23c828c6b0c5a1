// tensor_scan.cu -- tcgen05 prefilter for the batched brute-force cosine scan (S1, large B).
//
// The dense query x corpus product of configs C2/C5 genuinely is a contraction, so it runs
// on the 5th-gen tensor cores: fp16 copies of the L2-normalised rows ("shadow", built at
// append time) and of the normalised queries are multiplied with tcgen05.mma.kind::f16
// (fp32 accumulators in TMEM).  The result is an APPROXIMATE cosine a(q,x) with a rigorous
// error bound eps; it is used only to discard rows that provably
// cannot be in the exact top-k:
//     every epilogue thread owns one query (one TMEM lane) and keeps, in shared memory,
//     the k best approximate scores it has seen (A_k = the k-th).  k rows with a >= A_k
//     exist, so the exact k-th best score is >= A_k - eps, so a row can be in the exact
//     top-k only if a >= A_k - 2*eps.  Rows passing that test are appended to the query's
//     candidate list; A_k is shared between CTAs through a global atomicMax.
// The candidates are then re-scored by rerank_f32_kernel with the reference's exact
// arithmetic, so the final ids/scores are bit-identical to the exact scan.
//
// Kernel anatomy (one CTA per SM, 192 threads):
//   warp 0      TMA producer   cp.async.bulk.tensor.2d, SWIZZLE_128B, 4-stage mbarrier ring
//   warp 1      MMA issuer     tcgen05.mma cta_group::1 M128 x N256 x K16, commit -> mbarrier
//   warps 2..5  epilogue       tcgen05.ld 32x32b.x32 from a double-buffered TMEM accumulator
// Tile: 128 queries (A, K-major) x 256 corpus rows (B, K-major), K block = 64 halfs (128 B).
#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace cdb {

constexpr int TS_BLOCK_M = 128;
constexpr int TS_BLOCK_N = 256;
constexpr int TS_BLOCK_K = 64;  // fp16 elements = one 128-byte swizzle row
constexpr int TS_STAGES = 4;
constexpr int TS_THREADS = 192;
constexpr uint32_t TS_A_BYTES = TS_BLOCK_M * TS_BLOCK_K * 2;  // 16 KB
constexpr uint32_t TS_B_BYTES = TS_BLOCK_N * TS_BLOCK_K * 2;  // 32 KB
constexpr uint32_t TS_STAGE_BYTES = TS_A_BYTES + TS_B_BYTES;
constexpr uint32_t TS_TMEM_COLS = 512;  // 2 accumulator stages x 256 fp32 columns

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, both K-major, N >> 3, M >> 4
constexpr uint32_t TS_IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TS_BLOCK_N >> 3) << 17) | ((uint32_t)(TS_BLOCK_M >> 4) << 24);

struct TensorScanArgs {
    uint64_t n_rows;     // rows visible to this launch (the seeding pass sees a prefix)
    uint32_t n_queries;
    uint32_t k;
    uint32_t kblocks, mtiles, ntiles;
    float two_eps;
    uint32_t id_base;
    uint32_t emit;       // 0: seeding pass (only tightens gthr), 1: emit candidates
    int *ggm;            // [mtiles][64][128] shared class maxima (ordered int, class-major so a warp's 32 queries are
                         // contiguous), init f2ord(-inf)
    uint32_t *cand;      // [n_queries][cand_cap] global ids
    uint32_t *cand_cnt;  // [n_queries]
    uint32_t cand_cap;
    uint32_t *progress;  // [groups] tiles completed by the CTAs of a group (bounded-drift window), zeroed per launch
    uint32_t window;     // a CTA may run at most `window` tiles ahead of the slowest CTA of its group
    uint32_t refresh_mask;  // the shared bound is refreshed when (tile & mask) == mask (and for the first tiles)
    uint32_t has_deg;       // the corpus holds degenerate rows: kernel variant that keeps zero accumulators out of the bound
    // ---- integer form (KIND 1: exact scoring of u8 codes / sub-byte digits on kind::i8, tensor_scan_u8.cu)
    int metric;             // CDB_METRIC_COSINE or CDB_METRIC_DOT_PRODUCT
    const float *mags;      // [n_rows] stored magnitudes (cosine)
    const float *qmags;     // [n_queries]
    uint64_t *cand64;       // [n_queries][cand_cap] emitted selection keys (exact scores)
    uint32_t *err32;        // [n_queries] error bits
    float rel;              // threshold = bound * rel - two_eps (1 for the f16 form and for the dot product; 1 - 2^-18 for cosine)
};

constexpr uint32_t TS_LSTAGE = 32;  // thread-private candidate staging slots (shared memory), slot 0 = count
constexpr int TS_GROUPS = 64;       // at most 64 class maxima per query; the filter supports k <= TS_GROUPS

// rare path: one approximate score passed the threshold -> stage the row id, flush 32 at a time
__device__ __noinline__ void epi_emit(const TensorScanArgs &a, uint32_t qi, uint32_t *stage, uint32_t id) {
    uint32_t c = stage[0];
    stage[1 + c] = id;
    if (++c == TS_LSTAGE) {
        const uint32_t pos = atomicAdd(a.cand_cnt + qi, c);
        for (uint32_t i = 0; i < c; ++i)
            if (pos + i < a.cand_cap) a.cand[(size_t)qi * a.cand_cap + pos + i] = stage[1 + i];
        c = 0;
    }
    stage[0] = c;
}

// integer form: a value passed the filter -> its EXACT score (dp as f32, or dp / (|q| * |row|), cosine.rs:120, 223-235) becomes
// a selection key; 16 keys are staged per thread, slot 0 of the stage = count
constexpr uint32_t TS_KSTAGE = 16;
__device__ __noinline__ void epi_emit_key(const TensorScanArgs &a, uint32_t qi, uint64_t *stage, int dp, uint64_t row, float qmag) {
    float score = __int2float_rn(dp);   // u64 -> f32 of the reference; dp < 2^31 so the conversion is the same RNE
    if (a.metric == CDB_METRIC_COSINE) {
        const float denom = __fmul_rn(qmag, a.mags[row]);
        if (denom == 0.0f) return;      // Err(CalculationError): flagged per tile / per query in the epilogue, the row is skipped
        score = __fdiv_rn(score, denom);
    }
    uint32_t c = (uint32_t)stage[0];
    stage[1 + c] = make_key64(order_key(a.metric, __float_as_uint(score)), a.id_base + (uint32_t)row);
    if (++c == TS_KSTAGE) {
        const uint32_t pos = atomicAdd(a.cand_cnt + qi, c);
        for (uint32_t i = 0; i < c; ++i)
            if (pos + i < a.cand_cap) a.cand64[(size_t)qi * a.cand_cap + pos + i] = stage[1 + i];
        c = 0;
    }
    stage[0] = c;
}

// bitonic sort of NG register values, descending (fully unrolled: all indices are compile-time)
template <int NG>
__device__ __forceinline__ void sort_desc(float (&v)[NG]) {
#pragma unroll
    for (int size = 2; size <= NG; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int t = 0; t < NG / 2; ++t) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const float x = v[lo], y = v[hi];
                const float mx = fmaxf(x, y), mn = fminf(x, y);
                v[lo] = desc ? mx : mn;
                v[hi] = desc ? mn : mx;
            }
        }
    }
}

// HAS_DEG: the corpus holds degenerate rows (zero / non-finite / out-of-range norm).  Their shadow rows are all zero, so
// they score a == 0.0f exactly, but their exact score is NaN (worst) or +-inf -- the error bound does not cover them.
// They must not tighten the bound: accumulators that are exactly 0.0f are left out of the class maxima (dropping values
// from a maximum only loosens a lower bound, so a proper row that happens to score 0.0f is harmless).  Degenerate rows
// that are not all-zero are put on every candidate list by the host side (api.cu), all-zero rows are provably worst.
// CTAS = 2: two CTAs of a cluster (one TPC) work as a pair on query tiles 2p and 2p+1 against the same corpus tile:
// tcgen05.mma.cta_group::2 with M = 256 (each CTA's 128 queries stay in its own shared memory and TMEM) and N = 256 corpus
// rows of which each CTA stages only HALF (128 rows).  Per SM and k-block the shared-memory traffic drops from 96 KB
// (fill 48 + operand reads 48) to 64 KB (fill 32 + reads 32) against 542 tensor-pipe cycles of work -- the single-CTA form
// is bound by the 128 B/clk shared-memory port.  Barriers of the pair live in the leader (even) CTA:
// both CTAs' TMA loads complete on its full barrier, its MMA thread issues for both, commits multicast to both CTAs'
// empty / accumulator-full barriers, and both epilogues release the accumulator stage on its accumulator-empty barrier.
// KIND = 1: the same pipeline on tcgen05.mma.kind::i8 (u8 x u8 -> s32, exact): operand rows are 128 BYTES of K per stage
// (128 u8 elements instead of 64 halfs -- identical tile bytes, descriptors and swizzle), and the epilogue's values are
// exact integer dot products: the class-maximum bound works on f32(dp) (dot product) or on f32(dp) / |row| (cosine, the
// per-row reciprocals staged in shared memory per tile), the filter keeps a 2^-18 relative margin for the cosine form and
// none for the dot product, and whatever passes is scored EXACTLY (epi_emit_key) -- final selection keys, not candidates.
template <int STAGES, int NG, bool HAS_DEG, int CTAS, int KIND>
__global__ void __launch_bounds__(TS_THREADS, 1)
tensor_scan_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, TensorScanArgs a) {
    constexpr uint32_t B_ROWS = TS_BLOCK_N / CTAS;                          // corpus rows staged by this CTA
    constexpr uint32_t B_BYTES = B_ROWS * TS_BLOCK_K * 2;                   // 128 bytes of K per row in either form
    constexpr uint32_t STAGE_BYTES = TS_A_BYTES + B_BYTES;                  // 48 KB, resp. 32 KB per CTA of a pair
    constexpr uint32_t K_ELEMS = KIND == 1 ? 2 * TS_BLOCK_K : TS_BLOCK_K;   // elements per 128-byte row: 128 u8 / 64 halfs
    // instruction descriptor: D format (1 = F32, 2 = S32) | A/B formats (0 = F16 resp. UINT8) | N >> 3 | M >> 4
    constexpr uint32_t IDESC = ((KIND == 1 ? 2u : 1u) << 4) | ((uint32_t)(TS_BLOCK_N >> 3) << 17) | ((uint32_t)((TS_BLOCK_M * CTAS) >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *tiles = smem;
    uint32_t *lstage = reinterpret_cast<uint32_t *>(tiles + STAGES * STAGE_BYTES);           // [128][TS_LSTAGE+1]
    uint64_t *bars = reinterpret_cast<uint64_t *>(lstage + TS_BLOCK_M * (TS_LSTAGE + 2));     // 8-byte aligned
    uint64_t *full_bar = bars, *empty_bar = bars + STAGES, *tfull_bar = bars + 2 * STAGES, *tempty_bar = bars + 2 * STAGES + 2;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
    float *rinv_s = reinterpret_cast<float *>(tmem_ptr_smem + 4);           // KIND 1 cosine: [2][256] reciprocal row magnitudes
    uint32_t *zflag_s = reinterpret_cast<uint32_t *>(rinv_s + 2 * TS_BLOCK_N);   // [2] a valid row of the tile has |row| = 0

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = CTAS == 2 ? cluster_ctarank() : 0u;              // 0 = leader of the pair
    const bool leader = crank == 0;

    // work assignment: CTA c owns query tile c % mtiles and every G-th corpus tile
    // gridDim.x is a multiple of mtiles: the mtiles CTAs of group g walk the same corpus tiles g, g+G, ...
    // and are kept within `window` tiles of each other, so a corpus tile is fetched from HBM once and
    // served to the other query tiles from L2.  (Pairs: mtiles is even, CTAs 2j and 2j+1 of a group are one cluster.)
    const uint32_t mt = blockIdx.x % a.mtiles;
    const uint32_t g = blockIdx.x / a.mtiles;
    const uint32_t G = gridDim.x / a.mtiles;
    const uint32_t issuers = a.mtiles / CTAS;                               // MMA issuers per group (progress counter units)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull_bar[s]), 1); mbar_init(smem_u32(&tempty_bar[s]), 128 * CTAS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CTAS == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TS_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TS_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();               // barriers of both CTAs initialised before any remote arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer (one per CTA) =====================
        if (lane == 0) {
            uint32_t s = 0, phase = 0, t = 0;
            for (uint32_t nt = g; nt < a.ntiles; nt += G, ++t) {
                if (a.mtiles > 1 && t >= a.window) {
                    // bounded drift: every issuer of the group must have issued tile t - window
                    const uint32_t need = (t - a.window + 1) * issuers;
                    // bounded spin: if a peer CTA is not resident (GPU shared with another stream) we only lose locality
                    for (int spin = 0; spin < 20000 && *reinterpret_cast<volatile uint32_t *>(a.progress + g) < need; ++spin) __nanosleep(64);
                }
                for (uint32_t kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(smem_u32(&empty_bar[s]), phase ^ 1);          // own stage free (commit multicast reaches both CTAs)
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    if (CTAS == 2) {
                        if (leader) mbar_expect_tx(fb, 2 * STAGE_BYTES);     // the leader's barrier counts both CTAs' bytes
                        tma_load_2d_pair(sa, &map_q, fb, (int)(kb * K_ELEMS), (int)(mt * TS_BLOCK_M));
                        tma_load_2d_pair(sa + TS_A_BYTES, &map_x, fb, (int)(kb * K_ELEMS), (int)(nt * TS_BLOCK_N + crank * B_ROWS));
                    } else {
                        mbar_expect_tx(fb, STAGE_BYTES);
                        tma_load_2d(sa, &map_q, fb, (int)(kb * K_ELEMS), (int)(mt * TS_BLOCK_M));
                        tma_load_2d(sa + TS_A_BYTES, &map_x, fb, (int)(kb * K_ELEMS), (int)(nt * TS_BLOCK_N));
                    }
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (the leader issues for the pair) =====================
        if (lane == 0 && leader) {
            uint32_t s = 0, phase = 0, as = 0, aphase = 0;
            for (uint32_t nt = g; nt < a.ntiles; nt += G) {
                mbar_wait(smem_u32(&tempty_bar[as]), aphase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + as * TS_BLOCK_N;
                for (uint32_t kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(smem_u32(&full_bar[s]), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + TS_A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < TS_BLOCK_K / 16; ++kk) {  // +32 bytes along K = +2 in the encoded start address
                        if (KIND == 1) {
                            if (CTAS == 2) tcgen05_mma_i8_pair(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, IDESC, (kb | kk) != 0);
                            else tcgen05_mma_i8(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, IDESC, (kb | kk) != 0);
                        } else {
                            if (CTAS == 2) tcgen05_mma_f16_pair(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, IDESC, (kb | kk) != 0);
                            else tcgen05_mma_f16(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, IDESC, (kb | kk) != 0);
                        }
                    }
                    // frees the smem stage (of both CTAs) when these MMAs retire
                    if (CTAS == 2) tcgen05_commit_pair(smem_u32(&empty_bar[s])); else tcgen05_commit(smem_u32(&empty_bar[s]));
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
                // accumulator complete (in both CTAs' TMEM)
                if (CTAS == 2) tcgen05_commit_pair(smem_u32(&tfull_bar[as])); else tcgen05_commit(smem_u32(&tfull_bar[as]));
                if (a.mtiles > 1) atomicAdd(a.progress + g, 1u);  // all loads of this tile have landed in smem
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: threshold filter =====================
        const uint32_t lane_base = (uint32_t)(warp & 3) * 32;  // a warp may only touch TMEM lanes 32*(warp%4)..+31
        const uint32_t ql = lane_base + lane;                  // query within the tile == TMEM lane
        const uint32_t qi = mt * TS_BLOCK_M + ql;
        const bool qvalid = qi < a.n_queries;
        // Bound maintenance without a heap and without divergence: the thread keeps the maximum of each of NG
        // row classes (class = row mod NG; NG = 16/32/64 chosen from k).  The classes are disjoint row sets, so "the k-th largest class
        // maximum" is a valid lower bound of A_k; the maxima are shared between CTAs through atomicMax on
        // ggm[query][64] (class sets of different CTAs are disjoint too), which makes the bound GLOBAL: about as
        // tight as the exact running k-th best, at one FMNMX per value.
        uint32_t *stage = lstage + (size_t)ql * (TS_LSTAGE + 1);
        uint64_t *kstage = reinterpret_cast<uint64_t *>(lstage) + (size_t)ql * (TS_KSTAGE + 1);   // KIND 1: same bytes, 16 keys
        if (KIND == 1) kstage[0] = 0ull; else stage[0] = 0;
        const bool cosine = KIND == 1 && a.metric == CDB_METRIC_COSINE;
        const float qmag = (KIND == 1 && qvalid) ? a.qmags[qi] : 0.0f;
        uint32_t err = (cosine && qvalid && qmag == 0.0f && a.n_rows > 0) ? (uint32_t)CDB_ERRFLAG_CALCULATION : 0u;
        const uint32_t et = threadIdx.x - 64;                  // 0..127 within the epilogue warps
        float gm[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) gm[i] = -INFINITY;
        float bound = -INFINITY, thr = -INFINITY;
        int *ggm_q = a.ggm + (size_t)mt * NG * TS_BLOCK_M + ql;  // + class * 128
        if (qvalid) {  // what earlier launches (the seeding pass) and other CTAs already know
            float g2[NG];
#pragma unroll
            for (int i = 0; i < NG; ++i) g2[i] = ord2f(*reinterpret_cast<volatile int *>(ggm_q + (size_t)i * TS_BLOCK_M));
            sort_desc<NG>(g2);
#pragma unroll
            for (int i = 0; i < NG; ++i) if ((uint32_t)i == a.k - 1) bound = g2[i];
            thr = KIND == 1 ? (cosine ? bound * a.rel : bound) : bound - a.two_eps;
        }
        uint32_t as = 0, aphase = 0, t = 0;
        for (uint32_t nt = g; nt < a.ntiles; nt += G, ++t) {
            const uint64_t row0 = (uint64_t)nt * TS_BLOCK_N;
            const float *rs = rinv_s + as * TS_BLOCK_N;
            if (cosine) {
                // reciprocal magnitudes of this tile's rows -> shared memory (double buffered by accumulator stage); a zero
                // magnitude is an Err for every query that meets the row (cosine.rs:230-231)
                if (et == 0) zflag_s[as] = 0u;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (uint32_t c = et; c < (uint32_t)TS_BLOCK_N; c += 128) {
                    const bool valid = row0 + c < a.n_rows;
                    const float mg = valid ? a.mags[row0 + c] : 0.0f;
                    rinv_s[as * TS_BLOCK_N + c] = mg > 0.0f ? __frcp_rn(mg) : 0.0f;
                    if (valid && mg == 0.0f) zflag_s[as] = 1u;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (qvalid && zflag_s[as]) err |= (uint32_t)CDB_ERRFLAG_CALCULATION;
            }
            mbar_wait(smem_u32(&tfull_bar[as]), aphase);
            tcgen05_fence_after();
            const bool full_tile = row0 + TS_BLOCK_N <= a.n_rows;  // only the corpus' last tile can be partial (TMA zero fill)
#pragma unroll 1
            for (int c = 0; c < TS_BLOCK_N / 64; ++c) {
                uint32_t r0[32], r1[32];
                const uint32_t taddr = tmem_base + (lane_base << 16) + as * TS_BLOCK_N + c * 64;
                tmem_ld_32x32(taddr, r0);
                tmem_ld_32x32(taddr + 32, r1);
                tmem_ld_wait();
                uint32_t m0 = 0, m1 = 0;
                if (full_tile) {
                    // common case: fold the 64 values into the class maxima (max trees, FMNMX/FMNMX3) and test only the
                    // chunk maximum against the threshold; the per-value masks are built on the rare hit
                    float cm[NG];
#pragma unroll
                    for (int i = 0; i < NG; ++i) cm[i] = -INFINITY;
                    float zmax = -INFINITY;   // HAS_DEG: 0.0f if some accumulator of the chunk is exactly zero
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float v0, v1;
                        if (KIND == 1) {
                            // dot product: u8 x u8 sums are >= 0, and non-negative s32 bit patterns order like non-negative floats
                            // (FMNMX and the compare keep denormals), so the raw accumulator IS the filter value -- no I2F.
                            // cosine: f32(dp) / |row| through the staged reciprocals
                            if (cosine) {
                                v0 = __int2float_rn((int)r0[j]) * rs[c * 64 + j]; v1 = __int2float_rn((int)r1[j]) * rs[c * 64 + 32 + j];
                            } else {
                                v0 = __uint_as_float(r0[j]); v1 = __uint_as_float(r1[j]);
                            }
                        } else {
                            v0 = __uint_as_float(r0[j]); v1 = __uint_as_float(r1[j]);
                        }
                        if (HAS_DEG) {
                            zmax = (v0 == 0.0f || v1 == 0.0f) ? 0.0f : zmax;
                            v0 = v0 == 0.0f ? -INFINITY : v0;
                            v1 = v1 == 0.0f ? -INFINITY : v1;
                        }
                        cm[j & (NG - 1)] = fmaxf(cm[j & (NG - 1)], v0);
                        cm[(32 + j) & (NG - 1)] = fmaxf(cm[(32 + j) & (NG - 1)], v1);
                    }
                    float cmax = zmax;   // the emission test sees every value, the bound only the non-zero ones
#pragma unroll
                    for (int i = 0; i < NG; ++i) { cmax = fmaxf(cmax, cm[i]); gm[i] = fmaxf(gm[i], cm[i]); }
                    if (cmax >= thr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float v0, v1;
                            if (KIND == 1 && cosine) {
                                v0 = __int2float_rn((int)r0[j]) * rs[c * 64 + j]; v1 = __int2float_rn((int)r1[j]) * rs[c * 64 + 32 + j];
                            } else {
                                v0 = __uint_as_float(r0[j]); v1 = __uint_as_float(r1[j]);
                            }
                            m0 |= (v0 >= thr ? 1u : 0u) << j;
                            m1 |= (v1 >= thr ? 1u : 0u) << j;
                        }
                    }
                } else {
                    const uint64_t left = a.n_rows > row0 + c * 64 ? a.n_rows - (row0 + c * 64) : 0;  // valid columns in this chunk
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float v0, v1;
                        if (KIND == 1 && cosine) {
                            v0 = __int2float_rn((int)r0[j]) * rs[c * 64 + j]; v1 = __int2float_rn((int)r1[j]) * rs[c * 64 + 32 + j];
                        } else {
                            v0 = __uint_as_float(r0[j]); v1 = __uint_as_float(r1[j]);
                        }
                        const bool ok0 = (uint64_t)j < left, ok1 = (uint64_t)(32 + j) < left;
                        m0 |= ((ok0 && v0 >= thr) ? 1u : 0u) << j;
                        m1 |= ((ok1 && v1 >= thr) ? 1u : 0u) << j;
                        if (ok0 && !(HAS_DEG && v0 == 0.0f)) gm[j & (NG - 1)] = fmaxf(gm[j & (NG - 1)], v0);
                        if (ok1 && !(HAS_DEG && v1 == 0.0f)) gm[(32 + j) & (NG - 1)] = fmaxf(gm[(32 + j) & (NG - 1)], v1);
                    }
                }
                if (qvalid && a.emit && (m0 | m1)) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if ((m0 >> j) & 1u) {
                            if (KIND == 1) epi_emit_key(a, qi, kstage, (int)r0[j], row0 + c * 64 + j, qmag);
                            else epi_emit(a, qi, stage, a.id_base + (uint32_t)(row0 + c * 64 + j));
                        }
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if ((m1 >> j) & 1u) {
                            if (KIND == 1) epi_emit_key(a, qi, kstage, (int)r1[j], row0 + c * 64 + 32 + j, qmag);
                            else epi_emit(a, qi, stage, a.id_base + (uint32_t)(row0 + c * 64 + 32 + j));
                        }
                }
            }
            tcgen05_fence_before();
            // the accumulator stage is free: the MMA warp can run ahead (pairs: both CTAs' epilogues report to the leader)
            if (CTAS == 2) mbar_arrive_leader(smem_u32(&tempty_bar[as])); else mbar_arrive(smem_u32(&tempty_bar[as]));
            if (++as == 2) { as = 0; aphase ^= 1; }
            // refresh the shared bound: every tile early on (and in the seeding pass), then every 8th tile
            if (qvalid && (t < 8 || (t & a.refresh_mask) == a.refresh_mask || !a.emit)) {
                float g2[NG];
#pragma unroll
                for (int i = 0; i < NG; ++i) {
                    int *p = ggm_q + (size_t)i * TS_BLOCK_M;
                    const int mine = f2ord(gm[i]);
                    const int seen = *reinterpret_cast<volatile int *>(p);
                    if (mine > seen) atomicMax(p, mine);
                    g2[i] = ord2f(mine > seen ? mine : seen);
                }
                sort_desc<NG>(g2);
                float nb = -INFINITY;
#pragma unroll
                for (int i = 0; i < NG; ++i) if ((uint32_t)i == a.k - 1) nb = g2[i];
                if (nb > bound) { bound = nb; thr = KIND == 1 ? (cosine ? bound * a.rel : bound) : bound - a.two_eps; }
            }
        }
        if (qvalid) {  // final publish + flush of the staged candidates
#pragma unroll
            for (int i = 0; i < NG; ++i) atomicMax(ggm_q + (size_t)i * TS_BLOCK_M, f2ord(gm[i]));
            const uint32_t c = KIND == 1 ? (uint32_t)kstage[0] : stage[0];
            if (a.emit && c) {
                const uint32_t pos = atomicAdd(a.cand_cnt + qi, c);
                for (uint32_t i = 0; i < c; ++i)
                    if (pos + i < a.cand_cap) {
                        if (KIND == 1) a.cand64[(size_t)qi * a.cand_cap + pos + i] = kstage[1 + i];
                        else a.cand[(size_t)qi * a.cand_cap + pos + i] = stage[1 + i];
                    }
            }
            if (KIND == 1 && err) atomicOr(a.err32 + qi, err);
        }
    }
    tcgen05_fence_before();
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();   // pairs: no CTA leaves while its peer may still touch its barriers / smem
    if (warp == 1) {
        tcgen05_fence_after();
        if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TS_TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TS_TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ fp16 normalised copies
// x_hat = x / |x| (IEEE), rounded to fp16.  Rows whose norm is outside [1e-15, 1e15] (zero, underflowed, overflowed,
// inf/NaN) are DEGENERATE: the error bound of the prefilter does not cover them, their shadow row is all zero.
__host__ __device__ inline bool norm_is_proper(float m) { return m >= 1.0e-15f && m <= 1.0e15f; }

__global__ void normalize_f16_kernel(const float *__restrict__ raw, uint32_t pitch_elems, const float *__restrict__ mags,
                                     uint64_t n, uint32_t dim, __half *__restrict__ out, uint32_t out_pitch) {
    const uint32_t groups = (dim + 7) / 8;
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * groups) return;
    uint64_t row = gid / groups;
    uint32_t c0 = (uint32_t)(gid % groups) * 8;
    const float m = mags[row];
    const bool ok = norm_is_proper(m);
    for (uint32_t e = 0; e < 8 && c0 + e < dim; ++e) {
        float v = raw[row * pitch_elems + c0 + e];
        out[row * out_pitch + c0 + e] = __float2half_rn(ok ? __fdiv_rn(v, m) : 0.0f);
    }
}

cdb_status normalize_f16_device(const float *d_raw, uint32_t pitch_elems, const float *d_mags, uint64_t n, uint32_t dim,
                                void *d_out, uint32_t out_pitch_halfs, cudaStream_t s) {
    if (!n) return CDB_OK;
    uint64_t total = n * ((dim + 7) / 8);
    normalize_f16_kernel<<<(uint32_t)((total + 255) / 256), 256, 0, s>>>(d_raw, pitch_elems, d_mags, n, dim,
                                                                          reinterpret_cast<__half *>(d_out), out_pitch_halfs);
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

// Degenerate rows of an appended range.  deg = {all-zero rows, other degenerate rows, ids of the first TS_MAX_ODD of those}.
// An all-zero row scores 0/0 = NaN against every finite query: provably last.  Any other degenerate row (norm underflowed
// to 0 with non-zero elements -> +-inf, overflowed norm -> 0 or NaN, inf/NaN elements) can land anywhere in the exact
// ranking, so those rows are put on every query's candidate list and re-scored exactly.
__global__ void classify_rows_kernel(const float *__restrict__ raw, uint32_t pitch_elems, const float *__restrict__ mags,
                                     uint64_t n, uint32_t dim, uint32_t first_row, uint32_t *deg) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    if (norm_is_proper(mags[row])) return;
    bool allzero = true;
    for (uint32_t c = 0; c < dim && allzero; ++c) allzero = raw[row * pitch_elems + c] == 0.0f;
    if (allzero) { atomicAdd(deg, 1u); return; }
    const uint32_t pos = atomicAdd(deg + 1, 1u);
    if (pos < TS_MAX_ODD) deg[2 + pos] = first_row + (uint32_t)row;
}
cdb_status classify_rows_device(const float *d_raw, uint32_t pitch_elems, const float *d_mags, uint64_t n, uint32_t dim,
                                uint32_t first_row, uint32_t *d_deg, cudaStream_t s) {
    if (!n) return CDB_OK;
    classify_rows_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(d_raw, pitch_elems, d_mags, n, dim, first_row, d_deg);
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

// One launch prepares everything a prefilter search needs from the raw query batch (it replaces eleven memsets / small
// kernels whose launch gaps were a fixed ~0.1 ms per batch -- the limiter of strong scaling once the scan itself takes 1.6 ms):
//   q_raw   f32 copy in the index's pitched layout (operand of the exact re-rank), zero padded
//   q_mags  |q| = sqrt of the SEQUENTIAL fold of x*x (vector_store.rs:412), one thread, squares staged in shared memory
//   qh      fp16 copy of q / |q| (zero row for a degenerate norm and for the padding rows up to mtiles*128)
//   ggm     class maxima of the query := -inf;  cand_cnt := number of pre-seeded odd rows (+ their ids);  err := 0
//   progress counters of the seeding and of the main pass := 0
// One CTA of 128 threads per query slot (mtiles*128 slots).
__global__ void __launch_bounds__(128) prep_queries_kernel(const float *__restrict__ q, uint32_t nq, uint32_t dim, float *__restrict__ q_raw,
                                                           uint32_t raw_pitch_elems, float *__restrict__ q_mags, __half *__restrict__ qh,
                                                           uint32_t qh_pitch, int *__restrict__ ggm, const uint32_t *__restrict__ deg,
                                                           uint32_t has_deg, uint32_t id_base, uint32_t *__restrict__ cand,
                                                           uint32_t cand_cap, uint32_t *__restrict__ cand_cnt, uint32_t *__restrict__ err32,
                                                           uint8_t *__restrict__ err8, uint32_t *__restrict__ progress) {
    extern __shared__ float sq[];     // [dim] squares
    __shared__ float s_mag;
    const uint32_t slot = blockIdx.x, tid = threadIdx.x;
    if (slot == 0) for (uint32_t i = tid; i < 2048; i += 128) progress[i] = 0u;
    // class maxima: layout [tile][class][128]
    if (tid < (uint32_t)TS_GROUPS) ggm[((size_t)(slot >> 7) * TS_GROUPS + tid) * TS_BLOCK_M + (slot & 127u)] = (int)0x807FFFFF;   // f2ord(-inf)
    __half *hrow = qh + (size_t)slot * qh_pitch;
    if (slot >= nq) {
        for (uint32_t c = tid; c < qh_pitch; c += 128) hrow[c] = __float2half_rn(0.0f);
        return;
    }
    const float *src = q + (size_t)slot * dim;
    float *dst = q_raw + (size_t)slot * raw_pitch_elems;
    for (uint32_t c = tid; c < raw_pitch_elems; c += 128) {
        const float v = c < dim ? src[c] : 0.0f;
        dst[c] = v;
        if (c < dim) sq[c] = __fmul_rn(v, v);
    }
    __syncthreads();
    if (tid == 0) {
        float s = 0.0f;
        for (uint32_t c = 0; c < dim; ++c) s = __fadd_rn(s, sq[c]);
        const float m = __fsqrt_rn(s);
        s_mag = m;
        q_mags[slot] = m;
        err32[slot] = 0u;
        if (err8) err8[slot] = 0;
        uint32_t n_odd = 0;
        if (has_deg) {
            n_odd = min(min(deg[1], (uint32_t)TS_MAX_ODD), cand_cap);
            for (uint32_t i = 0; i < n_odd; ++i) cand[(size_t)slot * cand_cap + i] = id_base + deg[2 + i];
        }
        cand_cnt[slot] = n_odd;
    }
    __syncthreads();
    const float m = s_mag;
    const bool ok = norm_is_proper(m);
    for (uint32_t c = tid; c < qh_pitch; c += 128)
        hrow[c] = __float2half_rn((ok && c < dim) ? __fdiv_rn(src[c], m) : 0.0f);
}

cdb_status prep_queries_device(const float *d_q, uint32_t nq, uint32_t dim, float *d_q_raw, uint32_t raw_pitch_elems, float *d_q_mags,
                               void *d_qh, uint32_t qh_pitch_halfs, int *d_ggm, const uint32_t *d_deg, bool has_deg, uint32_t id_base,
                               uint32_t *d_cand, uint32_t cand_cap, uint32_t *d_cand_cnt, uint32_t *d_err32, uint8_t *d_err8,
                               uint32_t *d_progress, cudaStream_t s) {
    const uint32_t mtiles = (nq + TS_BLOCK_M - 1) / TS_BLOCK_M;
    prep_queries_kernel<<<mtiles * TS_BLOCK_M, 128, (size_t)dim * 4, s>>>(d_q, nq, dim, d_q_raw, raw_pitch_elems, d_q_mags,
                                                                          reinterpret_cast<__half *>(d_qh), qh_pitch_halfs, d_ggm, d_deg,
                                                                          has_deg ? 1u : 0u, id_base, d_cand, cand_cap, d_cand_cnt, d_err32,
                                                                          d_err8, d_progress);
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

// candidate lists start with the odd degenerate rows (global ids); cand_cnt[q] = n_odd
__global__ void seed_candidates_kernel(const uint32_t *__restrict__ deg, uint32_t id_base, uint32_t nq, uint32_t *cand,
                                       uint32_t cand_cap, uint32_t *cand_cnt) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint32_t n_odd = min(min(deg[1], (uint32_t)TS_MAX_ODD), cand_cap);
    for (uint32_t i = 0; i < n_odd; ++i) cand[(size_t)q * cand_cap + i] = id_base + deg[2 + i];
    cand_cnt[q] = n_odd;
}

__global__ void fill_i32_kernel(int *p, int v, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// Fallback selection: a query whose candidate list overflowed, or whose own norm is degenerate, is re-done by the exact
// scan.  qsel[0] = number of selected queries, qsel[1..] their indices.  flags[0] = that number (stats), flags[3] counts
// searches that needed any fallback.
__global__ void select_fallback_kernel(const uint32_t *__restrict__ cnt, uint32_t cap, const float *__restrict__ qmags,
                                       uint32_t n, uint32_t *qsel, uint32_t *flags) {
    __shared__ uint32_t total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const bool bad = cnt[i] > cap || (qmags && !norm_is_proper(qmags[i]));
        if (bad) qsel[1 + atomicAdd(&total, 1u)] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        qsel[0] = total;
        flags[0] = total;
        if (total) flags[3] += 1;
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

cdb_status make_tensor_map_2d(CUtensorMap *map, CUtensorMapDataType dtype, uint32_t elem_bytes, const void *base, uint64_t rows,
                              uint32_t cols, uint64_t pitch_bytes, uint32_t box_rows) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CDB_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return CDB_CUDA_ERROR; }
        fn = (PFN_encodeTiled)p;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {(cuuint64_t)pitch_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dtype, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return CDB_CUDA_ERROR; }
    return CDB_OK;
}
static cdb_status make_map_f16(CUtensorMap *map, const void *base, uint64_t rows, uint32_t dim, uint32_t pitch_halfs, uint32_t box_rows) {
    return make_tensor_map_2d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, rows, dim, (uint64_t)pitch_halfs * 2, box_rows);
}

static int tensor_scan_stages(uint32_t k) {
    if (k > (uint32_t)TS_GROUPS) return 0;  // the class-maximum bound needs k <= 64 disjoint classes
    return TS_STAGES;
}
constexpr int TS_STAGES_PAIR = 6;   // 32 KB per stage and CTA in the pair form: a deeper ring fits
static size_t tensor_scan_smem(int stages, uint32_t stage_bytes) {
    // tiles + per-thread staging + barriers + tmem pointer + (integer form) reciprocal magnitudes of two tiles + flags
    return 1024 + (size_t)stages * stage_bytes + (size_t)TS_BLOCK_M * (TS_LSTAGE + 2) * 4 + (2 * stages + 4) * 8 + 16 +
           2 * TS_BLOCK_N * 4 + 16;
}
size_t tensor_scan_smem_bytes(uint32_t k) {
    if (!tensor_scan_stages(k)) return (size_t)1 << 30;
    return tensor_scan_smem(TS_STAGES, TS_STAGE_BYTES);
}

template <int STAGES, int NG, bool HAS_DEG, int CTAS, int KIND>
static cdb_status launch_tensor_scan_gdc(const CUtensorMap &mq, const CUtensorMap &mx, const TensorScanArgs &a, uint32_t grid,
                                         cudaStream_t s) {
    auto kern = tensor_scan_kernel<STAGES, NG, HAS_DEG, CTAS, KIND>;
    const size_t smem = tensor_scan_smem(STAGES, TS_A_BYTES + TS_B_BYTES / CTAS);
    CDB_ALLOW_SMEM(kern, smem);
    if (CTAS == 1) {
        kern<<<grid, TS_THREADS, smem, s>>>(mq, mx, a);
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(TS_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        CDB_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, mq, mx, a));
    }
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

template <int NG, int CTAS, int KIND>
static cdb_status launch_tensor_scan_g(const CUtensorMap &mq, const CUtensorMap &mx, const TensorScanArgs &a, uint32_t grid, cudaStream_t s) {
    constexpr int ST = CTAS == 2 ? TS_STAGES_PAIR : TS_STAGES;
    if (KIND == 1) return launch_tensor_scan_gdc<ST, NG, false, CTAS, KIND>(mq, mx, a, grid, s);   // exact scores: no degenerate-row variant
    return a.has_deg ? launch_tensor_scan_gdc<ST, NG, true, CTAS, 0>(mq, mx, a, grid, s)
                     : launch_tensor_scan_gdc<ST, NG, false, CTAS, 0>(mq, mx, a, grid, s);
}

// fewer classes = cheaper bound refresh (the sort network grows as NG log^2 NG); the bound stays within ~1.5x of the
// exact k-th best as long as k is well below NG
template <int CTAS, int KIND>
static cdb_status launch_tensor_scan(const CUtensorMap &mq, const CUtensorMap &mx, const TensorScanArgs &a, uint32_t grid, cudaStream_t s) {
    if (a.k <= 12) return launch_tensor_scan_g<16, CTAS, KIND>(mq, mx, a, grid, s);
    if (a.k <= 28) return launch_tensor_scan_g<32, CTAS, KIND>(mq, mx, a, grid, s);
    return launch_tensor_scan_g<64, CTAS, KIND>(mq, mx, a, grid, s);
}

// seeding pass over a prefix of the corpus, then the main pass (both forms)
template <int KIND>
static cdb_status run_tensor_scan(const CUtensorMap &mq, const CUtensorMap &mx, TensorScanArgs a, bool pair, uint64_t n_rows,
                                  uint32_t *d_progress, int sm_count, cudaStream_t s) {
    cdb_status rc;
    // 1. seeding pass over a prefix of the corpus: a few CTAs per query tile, no emission, only the class maxima.
    //    Afterwards every CTA of the main pass starts from the k-th best of ~16K rows instead of from
    //    -inf, which removes the per-CTA warm-up bursts from the candidate lists.
    const uint32_t seed_tiles = std::min<uint32_t>(a.ntiles, 64);
    a.progress = d_progress;
    if (a.ntiles > 8) {
        TensorScanArgs sa = a;
        sa.emit = 0;
        sa.ntiles = seed_tiles;
        sa.n_rows = std::min<uint64_t>(n_rows, (uint64_t)seed_tiles * TS_BLOCK_N);
        // as many CTAs per query tile as the SMs allow, >= 2 corpus tiles each: the pass is a latency-bound prologue
        uint32_t per_m = std::max<uint32_t>(1, (uint32_t)sm_count / a.mtiles);
        per_m = std::min<uint32_t>(per_m, std::max<uint32_t>(1, seed_tiles / 2));
        const uint32_t sgrid = a.mtiles * per_m;
        rc = pair ? launch_tensor_scan<2, KIND>(mq, mx, sa, sgrid, s) : launch_tensor_scan<1, KIND>(mq, mx, sa, sgrid, s);
        if (rc) return rc;
    }
    // 2. main pass.  Every CTA should own several corpus tiles, so tiny corpora are not shredded over all SMs.
    a.emit = 1;
    uint64_t total_tiles = (uint64_t)a.mtiles * a.ntiles;
    uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sm_count, std::max<uint64_t>(1, total_tiles / 4));
    grid = std::max<uint32_t>(1, grid / a.mtiles) * a.mtiles;  // whole groups only (148 SMs, 8 query tiles -> 144 CTAs)
    if (grid > (uint32_t)sm_count) { set_error("tensor scan: more query tiles than SMs"); return CDB_INVALID_PARAMS; }
    a.progress = d_progress + 1024;   // the seeding pass used the first 1024 counters
    return pair ? launch_tensor_scan<2, KIND>(mq, mx, a, grid, s) : launch_tensor_scan<1, KIND>(mq, mx, a, grid, s);
}

static bool tensor_scan_pair_ok(uint32_t mtiles) {
    // CTA pairs (cta_group::2) need an even number of query tiles; CDB_TS_PAIR=0 forces the single-CTA form (A/B runs)
    static const bool pair_allowed = !(getenv("CDB_TS_PAIR") && atoi(getenv("CDB_TS_PAIR")) == 0);
    return pair_allowed && mtiles % 2 == 0;
}

// d_xh: fp16 normalised corpus [n_rows][pitch_halfs]; d_qh: fp16 normalised queries, padded with zero
// rows to a multiple of 128 [mtiles*128][pitch_halfs].
cdb_status tensor_scan_device(const void *d_xh, const void *d_qh, uint32_t pitch_halfs, uint64_t n_rows, uint32_t nq,
                              uint32_t dim, uint32_t k, float two_eps, uint32_t id_base, int *d_ggm, uint32_t *d_cand,
                              uint32_t *d_cand_cnt, uint32_t cand_cap, uint32_t *d_progress, const uint32_t *d_deg, bool has_deg,
                              int sm_count, cudaStream_t s, bool prepared) {
    TensorScanArgs a{};
    a.has_deg = has_deg ? 1u : 0u;
    a.progress = d_progress;
    a.window = 3;
    a.refresh_mask = 15;
    if (const char *e = getenv("CDB_TS_WINDOW")) a.window = (uint32_t)atoi(e);      // tuning knobs (bench experiments)
    if (const char *e = getenv("CDB_TS_REFRESH")) a.refresh_mask = (uint32_t)atoi(e);
    a.n_rows = n_rows;
    a.n_queries = nq;
    a.k = k;
    a.kblocks = (dim + TS_BLOCK_K - 1) / TS_BLOCK_K;
    a.mtiles = (nq + TS_BLOCK_M - 1) / TS_BLOCK_M;
    a.ntiles = (uint32_t)((n_rows + TS_BLOCK_N - 1) / TS_BLOCK_N);
    a.two_eps = two_eps;
    a.id_base = id_base;
    a.ggm = d_ggm;
    a.cand = d_cand;
    a.cand_cnt = d_cand_cnt;
    a.cand_cap = cand_cap;
    const int stages = tensor_scan_stages(k);
    if (!stages) { set_error("tensor scan: k too large for shared memory"); return CDB_INVALID_PARAMS; }
    const bool pair = tensor_scan_pair_ok(a.mtiles);
    CUtensorMap mq, mx;
    cdb_status rc;
    if ((rc = make_map_f16(&mq, d_qh, (uint64_t)a.mtiles * TS_BLOCK_M, dim, pitch_halfs, TS_BLOCK_M))) return rc;
    if ((rc = make_map_f16(&mx, d_xh, n_rows, dim, pitch_halfs, pair ? TS_BLOCK_N / 2 : TS_BLOCK_N))) return rc;
    if (!prepared) {   // prep_queries_kernel did all of this already
        fill_i32_kernel<<<(a.mtiles * TS_BLOCK_M * TS_GROUPS + 255) / 256, 256, 0, s>>>(d_ggm, (int)0x807FFFFF /* f2ord(-inf) */, a.mtiles * TS_BLOCK_M * TS_GROUPS);
        CDB_LAUNCH_CHECK();
        if (has_deg) {
            seed_candidates_kernel<<<(nq + 255) / 256, 256, 0, s>>>(d_deg, id_base, nq, d_cand, cand_cap, d_cand_cnt);
            CDB_LAUNCH_CHECK();
        } else {
            CDB_CUDA_TRY(cudaMemsetAsync(d_cand_cnt, 0, (size_t)nq * 4, s));
        }
        CDB_CUDA_TRY(cudaMemsetAsync(d_progress, 0, 8192, s));
    }

    a.rel = 1.0f;
    return run_tensor_scan<0>(mq, mx, a, pair, n_rows, d_progress, sm_count, s);
}

// Integer form (tensor_scan_u8.cu): d_x u8 operand rows [n_rows][pitch] (u8 codes or unpacked digits), d_q the query operand
// rows padded with zero rows to a multiple of 128.  Emits exact selection keys into d_cand64; the caller sorts them.
cdb_status tensor_scan_i8_device(const uint8_t *d_x, const uint8_t *d_q, uint32_t pitch, uint64_t n_rows, uint32_t nq, uint32_t dim,
                                 uint32_t k, int metric, const float *d_mags, const float *d_qmags, uint32_t id_base, int *d_ggm,
                                 uint64_t *d_cand64, uint32_t *d_cand_cnt, uint32_t cand_cap, uint32_t *d_err32, uint32_t *d_progress,
                                 int sm_count, cudaStream_t s) {
    TensorScanArgs a{};
    a.window = 3;
    a.refresh_mask = 15;
    a.n_rows = n_rows;
    a.n_queries = nq;
    a.k = k;
    a.kblocks = (dim + 2 * TS_BLOCK_K - 1) / (2 * TS_BLOCK_K);   // 128 u8 elements per 128-byte row
    a.mtiles = (nq + TS_BLOCK_M - 1) / TS_BLOCK_M;
    a.ntiles = (uint32_t)((n_rows + TS_BLOCK_N - 1) / TS_BLOCK_N);
    a.two_eps = 0.0f;
    a.rel = metric == CDB_METRIC_COSINE ? 1.0f - 1.0f / 262144.0f : 1.0f;
    a.id_base = id_base;
    a.ggm = d_ggm;
    a.cand64 = d_cand64;
    a.cand_cnt = d_cand_cnt;
    a.cand_cap = cand_cap;
    a.metric = metric;
    a.mags = d_mags;
    a.qmags = d_qmags;
    a.err32 = d_err32;
    if (!tensor_scan_stages(k)) { set_error("u8 tensor scan: k too large (class-maximum bound needs k <= 64)"); return CDB_INVALID_PARAMS; }
    if ((uint32_t)sm_count < a.mtiles) { set_error("u8 tensor scan: more query tiles than SMs"); return CDB_INVALID_PARAMS; }
    const bool pair = tensor_scan_pair_ok(a.mtiles);
    CUtensorMap mq, mx;
    cdb_status rc;
    if ((rc = make_tensor_map_2d(&mq, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d_q, (uint64_t)a.mtiles * TS_BLOCK_M, dim, pitch, TS_BLOCK_M))) return rc;
    if ((rc = make_tensor_map_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d_x, n_rows, dim, pitch, pair ? TS_BLOCK_N / 2 : TS_BLOCK_N))) return rc;
    fill_i32_kernel<<<(a.mtiles * TS_BLOCK_M * TS_GROUPS + 255) / 256, 256, 0, s>>>(d_ggm, (int)0x807FFFFF /* f2ord(-inf) */, a.mtiles * TS_BLOCK_M * TS_GROUPS);
    CDB_LAUNCH_CHECK();
    CDB_CUDA_TRY(cudaMemsetAsync(d_cand_cnt, 0, (size_t)nq * 4, s));
    CDB_CUDA_TRY(cudaMemsetAsync(d_progress, 0, 8192, s));
    return run_tensor_scan<1>(mq, mx, a, pair, n_rows, d_progress, sm_count, s);
}

cdb_status select_fallback_device(const uint32_t *d_cnt, uint32_t cap, const float *d_qmags, uint32_t n, uint32_t *d_qsel,
                                  uint32_t *d_flags, cudaStream_t s) {
    select_fallback_kernel<<<1, 1024, 0, s>>>(d_cnt, cap, d_qmags, n, d_qsel, d_flags);
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

}  // namespace cdb
