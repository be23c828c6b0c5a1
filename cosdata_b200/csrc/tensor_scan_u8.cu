// tensor_scan_u8.cu -- EXACT integer scoring of quantized codes on the tensor cores (S1, BRUTE_CODES,
// configs C4 / the reference's default `UnsignedByte` storage).
//
// dot_product_u8 (src/models/dot_product/x86_64.rs:22-66) and the bit-plane products
// dot_product_binary/quaternary/octal (src/models/dot_product.rs:21-90) are integer contractions
//   sum_i a_i * b_i   with a_i, b_i in 0..255 (u8) or the sub-byte digits 0..1 / 0..3 / 0..7,
// so tcgen05.mma.kind::i8 (u8 x u8 -> s32 in TMEM) computes them bit-exactly: max 255*255*D < 2^31.
// u8 storage is consumed in place; sub-byte storage through an unpacked digit copy (digit =
// plane0 + 2*plane1 + 4*plane2, exactly the weights the reference kernels apply).
//
// Same anatomy as tensor_scan.cu (TMA producer warp, MMA warp, 4 epilogue warps, 2 TMEM
// accumulator stages, bounded-drift CTA groups), but the epilogue produces FINAL scores:
//   DotProduct:  score = dp as f32                       (dotproduct.rs:28,55-60)
//   Cosine:      score = (dp as f32) / (|q| * |row|)     (cosine.rs:120,223-235), 0 denominator -> Err
// Each epilogue thread owns one query, keeps the k best scores seen (A_k) and emits every row with
// score >= A_k as a 64-bit selection key; a final sort picks the top-k.  Most values are rejected
// by one int->float convert, one multiply and one compare (no division).
#include "kernels.h"
#include "tc_common.cuh"

namespace cdb {

constexpr int TU_BLOCK_M = 128;
constexpr int TU_BLOCK_N = 256;
constexpr int TU_BLOCK_K = 128;  // u8 elements = one 128-byte swizzle row
constexpr int TU_THREADS = 192;
constexpr uint32_t TU_A_BYTES = TU_BLOCK_M * TU_BLOCK_K;
constexpr uint32_t TU_B_BYTES = TU_BLOCK_N * TU_BLOCK_K;
constexpr uint32_t TU_STAGE_BYTES = TU_A_BYTES + TU_B_BYTES;
constexpr uint32_t TU_TMEM_COLS = 512;
constexpr uint32_t TU_LSTAGE = 16;  // staged keys per thread
// D = S32 (c_format 2), A = B = UINT8 (format 0), K-major, N >> 3, M >> 4
constexpr uint32_t TU_IDESC = (2u << 4) | ((uint32_t)(TU_BLOCK_N >> 3) << 17) | ((uint32_t)(TU_BLOCK_M >> 4) << 24);

__device__ __forceinline__ void tcgen05_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct TensorU8Args {
    uint64_t n_rows;
    uint32_t n_queries, k, kblocks, mtiles, ntiles;
    uint32_t id_base, emit;
    int metric;           // CDB_METRIC_COSINE or CDB_METRIC_DOT_PRODUCT
    const float *mags;    // [n_rows] stored magnitudes (cosine)
    const float *qmags;   // [n_queries]
    int *gthr;            // [n_queries] ordered-int lower bound of A_k
    uint64_t *cand;       // [n_queries][cand_cap] selection keys
    uint32_t *cand_cnt;
    uint32_t cand_cap;
    uint32_t *err32;      // [n_queries]
    uint32_t *progress;
    uint32_t window;
};

struct EpiU8 {
    float *heap;
    uint64_t *stage;
    uint32_t hcnt, scnt;
    float local_min, bound;
};

__device__ __forceinline__ void epi8_flush(const TensorU8Args &a, EpiU8 &st, uint32_t qi) {
    if (!st.scnt) return;
    const uint32_t pos = atomicAdd(a.cand_cnt + qi, st.scnt);
    for (uint32_t i = 0; i < st.scnt; ++i)
        if (pos + i < a.cand_cap) a.cand[(size_t)qi * a.cand_cap + pos + i] = st.stage[i];
    st.scnt = 0;
}

// slow path: exact score of one (query,row) pair that survived the cheap test
__device__ __noinline__ void epi8_accept(const TensorU8Args &a, EpiU8 &st, uint32_t qi, int dp, uint64_t row, float qmag) {
    if (row >= a.n_rows) return;
    float score = __int2float_rn(dp);  // u64 -> f32 of the reference; dp < 2^31 so the conversion is the same RNE
    if (a.metric == CDB_METRIC_COSINE) {
        const float denom = __fmul_rn(qmag, a.mags[row]);
        if (denom == 0.0f) { atomicOr(a.err32 + qi, (uint32_t)CDB_ERRFLAG_CALCULATION); return; }  // cosine.rs:230-231
        score = __fdiv_rn(score, denom);
    }
    if (!(score >= st.bound)) return;  // exact test (ties with A_k are kept)
    if (a.emit) {
        st.stage[st.scnt++] = make_key64(order_key(a.metric, __float_as_uint(score)), a.id_base + (uint32_t)row);
        if (st.scnt == TU_LSTAGE) epi8_flush(a, st, qi);
    }
    if (st.hcnt < a.k || score > st.local_min) {
        float *h = st.heap;
        if (st.hcnt < a.k) {
            uint32_t i = st.hcnt++;
            h[i] = score;
            while (i > 0) {
                uint32_t p = (i - 1) >> 1;
                if (h[p] <= h[i]) break;
                float t = h[p]; h[p] = h[i]; h[i] = t;
                i = p;
            }
        } else {
            h[0] = score;
            uint32_t i = 0;
            for (;;) {
                uint32_t l = 2 * i + 1, r = l + 1, m = i;
                if (l < a.k && h[l] < h[m]) m = l;
                if (r < a.k && h[r] < h[m]) m = r;
                if (m == i) break;
                float t = h[m]; h[m] = h[i]; h[i] = t;
                i = m;
            }
        }
        if (st.hcnt == a.k) {
            st.local_min = h[0];
            atomicMax(a.gthr + qi, f2ord(st.local_min));
            if (st.local_min > st.bound) st.bound = st.local_min;
        }
    }
}

template <int STAGES>
__global__ void __launch_bounds__(TU_THREADS, 1)
tensor_scan_u8_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, TensorU8Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *tiles = smem;
    uint64_t *lstage = reinterpret_cast<uint64_t *>(tiles + STAGES * TU_STAGE_BYTES);  // [128][TU_LSTAGE]
    float *mag_s = reinterpret_cast<float *>(lstage + TU_BLOCK_M * TU_LSTAGE);          // [2][256]
    float *heaps = mag_s + 2 * TU_BLOCK_N;                                             // [128][k]
    uint64_t *bars = reinterpret_cast<uint64_t *>(heaps + (size_t)TU_BLOCK_M * a.k);
    uint64_t *full_bar = bars, *empty_bar = bars + STAGES, *tfull_bar = bars + 2 * STAGES, *tempty_bar = bars + 2 * STAGES + 2;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t mt = blockIdx.x % a.mtiles;
    const uint32_t g = blockIdx.x / a.mtiles;
    const uint32_t G = gridDim.x / a.mtiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull_bar[s]), 1); mbar_init(smem_u32(&tempty_bar[s]), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TU_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t s = 0, phase = 0, t = 0;
            for (uint32_t nt = g; nt < a.ntiles; nt += G, ++t) {
                if (a.mtiles > 1 && t >= a.window) {
                    const uint32_t need = (t - a.window + 1) * a.mtiles;
                    for (int spin = 0; spin < 20000 && *reinterpret_cast<volatile uint32_t *>(a.progress + g) < need; ++spin) __nanosleep(64);
                }
                for (uint32_t kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(smem_u32(&empty_bar[s]), phase ^ 1);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, TU_STAGE_BYTES);
                    const uint32_t sa = smem_u32(tiles + (size_t)s * TU_STAGE_BYTES);
                    tma_load_2d(sa, &map_q, fb, (int)(kb * TU_BLOCK_K), (int)(mt * TU_BLOCK_M));
                    tma_load_2d(sa + TU_A_BYTES, &map_x, fb, (int)(kb * TU_BLOCK_K), (int)(nt * TU_BLOCK_N));
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t s = 0, phase = 0, as = 0, aphase = 0;
            for (uint32_t nt = g; nt < a.ntiles; nt += G) {
                mbar_wait(smem_u32(&tempty_bar[as]), aphase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + as * TU_BLOCK_N;
                for (uint32_t kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(smem_u32(&full_bar[s]), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(tiles + (size_t)s * TU_STAGE_BYTES);
                    const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + TU_A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < TU_BLOCK_K / 32; ++kk)  // K = 32 bytes per instruction
                        tcgen05_mma_i8(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, TU_IDESC, (kb | kk) != 0);
                    tcgen05_commit(smem_u32(&empty_bar[s]));
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
                tcgen05_commit(smem_u32(&tfull_bar[as]));
                if (a.mtiles > 1) atomicAdd(a.progress + g, 1u);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        const uint32_t lane_base = (uint32_t)(warp & 3) * 32;
        const uint32_t ql = lane_base + lane;
        const uint32_t qi = mt * TU_BLOCK_M + ql;
        const bool qvalid = qi < a.n_queries;
        const float qmag = qvalid ? a.qmags[qi] : 0.0f;
        const bool cosine = a.metric == CDB_METRIC_COSINE;
        EpiU8 st;
        st.heap = heaps + (size_t)ql * a.k;
        st.stage = lstage + (size_t)ql * TU_LSTAGE;
        st.hcnt = 0;
        st.scnt = 0;
        st.local_min = -INFINITY;
        st.bound = -INFINITY;
        const uint32_t et = threadIdx.x - 64;  // 0..127 within the epilogue warps
        uint32_t as = 0, aphase = 0;
        for (uint32_t nt = g; nt < a.ntiles; nt += G) {
            const uint64_t row0 = (uint64_t)nt * TU_BLOCK_N;
            int gnow = (int)0x807FFFFF;
            if (qvalid) gnow = *reinterpret_cast<volatile int *>(a.gthr + qi);
            float *ms = mag_s + as * TU_BLOCK_N;
            if (cosine) {  // stored magnitudes of this tile's rows -> shared memory (double buffered by accumulator stage)
                for (uint32_t c = et; c < (uint32_t)TU_BLOCK_N; c += 128) ms[c] = row0 + c < a.n_rows ? a.mags[row0 + c] : 0.0f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(smem_u32(&tfull_bar[as]), aphase);
            tcgen05_fence_after();
            {
                const float gb = ord2f(gnow);
                if (gb > st.bound) st.bound = gb;
            }
#pragma unroll 1
            for (int c = 0; c < TU_BLOCK_N / 64; ++c) {
                uint32_t r0[32], r1[32];
                const uint32_t taddr = tmem_base + (lane_base << 16) + as * TU_BLOCK_N + c * 64;
                tmem_ld_32x32(taddr, r0);
                tmem_ld_32x32(taddr + 32, r1);
                tmem_ld_wait();
                // cheap reject: score >= bound  <=>  dp >= bound * denom (relaxed by 2^-20 so rounding can only over-accept)
                const float bound = st.bound;
                const float tq = cosine ? __fmul_rn(__fmul_rn(bound, qmag), 0.99999905f) : bound;
                const bool all = !(bound > 0.0f);  // warm-up (or nothing known yet): take the slow path for everything
                uint32_t m0 = 0, m1 = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float rhs0 = cosine ? tq * ms[c * 64 + j] : tq;
                    const float rhs1 = cosine ? tq * ms[c * 64 + 32 + j] : tq;
                    m0 |= ((all || __int2float_rn((int)r0[j]) >= rhs0) ? 1u : 0u) << j;
                    m1 |= ((all || __int2float_rn((int)r1[j]) >= rhs1) ? 1u : 0u) << j;
                }
                if (qvalid && (m0 | m1)) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if ((m0 >> j) & 1u) epi8_accept(a, st, qi, (int)r0[j], row0 + c * 64 + j, qmag);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if ((m1 >> j) & 1u) epi8_accept(a, st, qi, (int)r1[j], row0 + c * 64 + 32 + j, qmag);
                }
            }
            tcgen05_fence_before();
            mbar_arrive(smem_u32(&tempty_bar[as]));
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (qvalid && a.emit) epi8_flush(a, st, qi);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TU_TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ final selection: sort the emitted keys
__global__ void __launch_bounds__(256) select_keys_kernel(const uint64_t *__restrict__ cand, const uint32_t *__restrict__ cand_cnt,
                                                          uint32_t cap, int metric, uint32_t k, uint32_t *__restrict__ ids,
                                                          float *__restrict__ scores, uint32_t *__restrict__ counts,
                                                          uint64_t *__restrict__ out_keys) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem);
    const uint32_t q = blockIdx.x, n = min(cand_cnt[q], cap);
    uint32_t P = 1;
    while (P < n) P <<= 1;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) keys[i] = i < n ? cand[(size_t)q * cap + i] : 0ull;
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) {
        ids[(size_t)q * k + j] = CDB_INVALID_ID;
        scores[(size_t)q * k + j] = 0.0f;
        if (out_keys) out_keys[(size_t)q * k + j] = 0ull;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= P; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < P / 2; t += blockDim.x) {
                const uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const uint64_t x = keys[lo], y = keys[hi];
                if ((x < y) == desc) { keys[lo] = y; keys[hi] = x; }
            }
            __syncthreads();
        }
    const uint32_t m = min(n, k);
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        ids[(size_t)q * k + i] = key64_id(keys[i]);
        scores[(size_t)q * k + i] = __uint_as_float(key_to_bits(metric, (uint32_t)(keys[i] >> 32)));
        if (out_keys) out_keys[(size_t)q * k + i] = keys[i];
    }
    if (threadIdx.x == 0 && counts) counts[q] = m;
}

// ------------------------------------------------------------------ sub-byte planes -> u8 digits
// digit = plane0 + 2*plane1 + 4*plane2: the weights dot_product_{binary,quaternary,octal} apply to the planes
__global__ void unpack_digits_kernel(const uint8_t *__restrict__ codes, uint32_t row_pitch, uint64_t n, uint32_t dim, int res,
                                     uint8_t *__restrict__ out, uint32_t out_pitch) {
    const uint32_t nb = plane_bytes(dim), pp = plane_pitch(dim);
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * nb) return;
    const uint64_t row = gid / nb;
    const uint32_t b = (uint32_t)(gid % nb);
    uint32_t pl[3] = {0, 0, 0};
    for (int p = 0; p < res; ++p) pl[p] = codes[row * row_pitch + (size_t)p * pp + b];
    uint8_t d[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = (uint8_t)(((pl[0] >> e) & 1u) | (((pl[1] >> e) & 1u) << 1) | (((pl[2] >> e) & 1u) << 2));
    uint8_t *o = out + row * out_pitch + (size_t)b * 8;
    for (int e = 0; e < 8 && b * 8 + e < dim; ++e) o[e] = d[e];
}

cdb_status unpack_digits_device(const uint8_t *d_codes, uint32_t row_pitch, uint64_t n, uint32_t dim, int res, uint8_t *d_out,
                                uint32_t out_pitch, cudaStream_t s) {
    if (!n) return CDB_OK;
    uint64_t total = n * plane_bytes(dim);
    unpack_digits_kernel<<<(uint32_t)((total + 255) / 256), 256, 0, s>>>(d_codes, row_pitch, n, dim, res, d_out, out_pitch);
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

// ------------------------------------------------------------------ host
static int tensor_u8_stages(uint32_t k) {
    auto bytes = [&](int stages) {
        return 1024 + (size_t)stages * TU_STAGE_BYTES + (size_t)TU_BLOCK_M * TU_LSTAGE * 8 + 2 * TU_BLOCK_N * 4 +
               (size_t)TU_BLOCK_M * k * 4 + (2 * stages + 4) * 8 + 16;
    };
    if (bytes(4) <= 227 * 1024) return 4;
    if (bytes(3) <= 227 * 1024) return 3;
    return 0;
}
size_t tensor_u8_smem_bytes(uint32_t k) {
    int st = tensor_u8_stages(k);
    if (!st) return (size_t)1 << 30;
    return 1024 + (size_t)st * TU_STAGE_BYTES + (size_t)TU_BLOCK_M * TU_LSTAGE * 8 + 2 * TU_BLOCK_N * 4 + (size_t)TU_BLOCK_M * k * 4 +
           (2 * st + 4) * 8 + 16;
}

__global__ void fill_i32_kernel2(int *p, int v, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <int STAGES>
static cdb_status launch_u8(const CUtensorMap &mq, const CUtensorMap &mx, const TensorU8Args &a, uint32_t grid, size_t smem, cudaStream_t s) {
    auto kern = tensor_scan_u8_kernel<STAGES>;
    CDB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, TU_THREADS, smem, s>>>(mq, mx, a);
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

// d_x: u8 operand rows [n_rows][pitch] (u8 codes or unpacked digits); d_q: query operand rows padded with zero
// rows to a multiple of 128.  Writes ids/scores/counts; *overflow flag* = d_flag[0] += #queries whose list overflowed.
cdb_status tensor_u8_scan_device(const uint8_t *d_x, const uint8_t *d_q, uint32_t pitch, uint64_t n_rows, uint32_t nq, uint32_t dim,
                                 uint32_t k, int metric, const float *d_mags, const float *d_qmags, uint32_t id_base, int *d_gthr,
                                 uint64_t *d_cand, uint32_t *d_cand_cnt, uint32_t cand_cap, uint32_t *d_err32, uint32_t *d_progress,
                                 uint32_t *d_ids, float *d_scores, uint32_t *d_counts, int sm_count, cudaStream_t s,
                                 uint64_t *d_out_keys) {
    TensorU8Args a{};
    a.n_rows = n_rows;
    a.n_queries = nq;
    a.k = k;
    a.kblocks = (dim + TU_BLOCK_K - 1) / TU_BLOCK_K;
    a.mtiles = (nq + TU_BLOCK_M - 1) / TU_BLOCK_M;
    a.ntiles = (uint32_t)((n_rows + TU_BLOCK_N - 1) / TU_BLOCK_N);
    a.id_base = id_base;
    a.metric = metric;
    a.mags = d_mags;
    a.qmags = d_qmags;
    a.gthr = d_gthr;
    a.cand = d_cand;
    a.cand_cnt = d_cand_cnt;
    a.cand_cap = cand_cap;
    a.err32 = d_err32;
    a.progress = d_progress;
    a.window = 3;
    const int stages = tensor_u8_stages(k);
    if (!stages) { set_error("u8 tensor scan: k too large for shared memory"); return CDB_INVALID_PARAMS; }
    const size_t smem = tensor_u8_smem_bytes(k);
    CUtensorMap mq, mx;
    cdb_status rc;
    if ((rc = make_tensor_map_2d(&mq, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d_q, (uint64_t)a.mtiles * TU_BLOCK_M, dim, pitch, TU_BLOCK_M))) return rc;
    if ((rc = make_tensor_map_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d_x, n_rows, dim, pitch, TU_BLOCK_N))) return rc;
    fill_i32_kernel2<<<(nq + 255) / 256, 256, 0, s>>>(d_gthr, (int)0x807FFFFF, nq);
    CDB_LAUNCH_CHECK();
    CDB_CUDA_TRY(cudaMemsetAsync(d_cand_cnt, 0, (size_t)nq * 4, s));
    if ((uint32_t)sm_count < a.mtiles) { set_error("u8 tensor scan: more query tiles than SMs"); return CDB_INVALID_PARAMS; }
    const uint32_t seed_tiles = std::min<uint32_t>(a.ntiles, 64);
    if (a.ntiles > 8) {
        TensorU8Args sa = a;
        sa.emit = 0;
        sa.ntiles = seed_tiles;
        sa.n_rows = std::min<uint64_t>(n_rows, (uint64_t)seed_tiles * TU_BLOCK_N);
        uint32_t per_m = std::max<uint32_t>(1, std::min<uint32_t>(8, (uint32_t)sm_count / a.mtiles));
        per_m = std::min<uint32_t>(per_m, std::max<uint32_t>(1, seed_tiles / 8));
        CDB_CUDA_TRY(cudaMemsetAsync(d_progress, 0, 4096, s));
        rc = stages == 4 ? launch_u8<4>(mq, mx, sa, a.mtiles * per_m, smem, s) : launch_u8<3>(mq, mx, sa, a.mtiles * per_m, smem, s);
        if (rc) return rc;
    }
    a.emit = 1;
    uint64_t total_tiles = (uint64_t)a.mtiles * a.ntiles;
    uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sm_count, std::max<uint64_t>(1, total_tiles / 4));
    grid = std::max<uint32_t>(1, grid / a.mtiles) * a.mtiles;
    CDB_CUDA_TRY(cudaMemsetAsync(d_progress, 0, 4096, s));
    rc = stages == 4 ? launch_u8<4>(mq, mx, a, grid, smem, s) : launch_u8<3>(mq, mx, a, grid, smem, s);
    if (rc) return rc;
    uint32_t P = 1;
    while (P < cand_cap) P <<= 1;
    const size_t ssm = (size_t)P * 8;
    CDB_CUDA_TRY(cudaFuncSetAttribute(select_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
    select_keys_kernel<<<nq, 256, ssm, s>>>(d_cand, d_cand_cnt, cand_cap, metric, k, d_ids, d_scores, d_counts, d_out_keys);
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

}  // namespace cdb
