// K1: brute-force tiled k-nearest-neighbour search with a fused IDW epilogue.
//
// Replaces scipy.spatial.cKDTree(points).query(targets, k, workers=-1)
// (reference src/climatrix/reconstruct/idw.py:155-158) and, with the epilogue,
// idw.py:163-180.
//
// Scheme (see DESIGN.md "K1"):
//   * every (target, sample) pair is evaluated on the FP32 pipe in the expanded
//     form  t = |s|^2 - 2 q.s  (2 FFMA + 1 compare per pair) in coordinates
//     centred on the target tile, against a per-target threshold held in a
//     register;
//   * samples are staged through shared memory by TMA bulk copies
//     (cp.async.bulk + mbarrier), re-centred and converted to fp32 per tile;
//   * pairs that pass the (conservative, error-bounded) fp32 filter are pushed to
//     a per-target candidate buffer; when a buffer fills, the warp re-ranks it
//     cooperatively in exact, non-contracted float64 (bit-identical to cKDTree's
//     dx*dx+dy*dy) with a bitonic sort/merge over warp shuffles and tightens the
//     threshold.  Ties break by sample index, so results are deterministic;
//   * the final re-rank leaves the k neighbours in registers, where the IDW
//     weights and the weighted mean are evaluated before anything is stored.
#include <limits.h>

#include "common.cuh"

namespace cmx {
namespace {

constexpr int WARPS = 4;
constexpr int THREADS = WARPS * 32;
constexpr int TS = 2048;             // samples per shared-memory stage
constexpr int CHK = 32;              // samples between candidate-buffer checks
constexpr int CAP = 64;              // candidate slots per target
constexpr int FLUSH_AT = CAP - CHK;  // re-rank a target once it holds more than this
constexpr int MAX_CTAS_PER_SM = 2;
constexpr int MAX_TQ = WARPS * 8 * 32;

#define CMX_INF __longlong_as_double(0x7ff0000000000000LL)
#define CMX_INF_F __int_as_float(0x7f800000)

struct KnnArgs {
    KnnJob j;
    long long M;
    int tiles_c;    // grid mode: tiles per row of tiles
    int num_tiles;
    int aligned;    // sample arrays 16-byte aligned -> TMA bulk staging
    unsigned* cand; // [ctas][TQ][CAP] candidate sample indices
    double* list_d; // [ctas][TQ][KP]  running top-KP squared distances (ascending)
    int* list_i;    // [ctas][TQ][KP]  and their sample indices
};

__device__ __forceinline__ bool lex_less(double da, int ia, double db, int ib) {
    return da < db || (da == db && ia < ib);
}

// ascending bitonic sort of one (d, i) element per lane
__device__ __forceinline__ void warp_sort32(double& d, int& i, int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            const double pd = __shfl_xor_sync(kFull, d, j);
            const int pi = __shfl_xor_sync(kFull, i, j);
            const bool up = (lane & k2) == 0;
            const bool lower = (lane & j) == 0;
            const bool p_less = lex_less(pd, pi, d, i);
            const bool take = (lower == up) ? p_less : !p_less;
            if (take) {
                d = pd;
                i = pi;
            }
        }
    }
}

// list: KP = 32*E elements, element e*32+lane in (ld[e], li[e]), ascending.
// (cd, ci): 32 sorted candidates, one per lane.  Keeps the KP smallest, sorted.
template <int E>
__device__ __forceinline__ void merge_sorted32(double (&ld)[E], int (&li)[E], double cd, int ci, int lane) {
    // min(list[i], cand_reversed[i]) over the top 32 slots -> bitonic sequence of the KP smallest
    const double rd = __shfl_sync(kFull, cd, 31 - lane);
    const int ri = __shfl_sync(kFull, ci, 31 - lane);
    if (lex_less(rd, ri, ld[E - 1], li[E - 1])) {
        ld[E - 1] = rd;
        li[E - 1] = ri;
    }
#pragma unroll
    for (int j = E / 2; j >= 1; j >>= 1) {  // strides of j*32 elements stay inside a lane
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if ((e & j) == 0 && e + j < E) {
                if (lex_less(ld[e + j], li[e + j], ld[e], li[e])) {
                    const double td = ld[e];
                    const int ti = li[e];
                    ld[e] = ld[e + j];
                    li[e] = li[e + j];
                    ld[e + j] = td;
                    li[e + j] = ti;
                }
            }
        }
    }
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const double pd = __shfl_xor_sync(kFull, ld[e], j);
            const int pi = __shfl_xor_sync(kFull, li[e], j);
            const bool lower = (lane & j) == 0;
            const bool p_less = lex_less(pd, pi, ld[e], li[e]);
            if (lower ? p_less : !p_less) {
                ld[e] = pd;
                li[e] = pi;
            }
        }
    }
}

// Conservative fp32 filter threshold for "exact d2 <= d2k" in t-space (t = d2 - |q|^2).
// Error model (DESIGN.md "filter margin"): coordinates are rounded to fp32 after
// centring (relative 2^-24 each), S = fma(b,b,a*a), t = fma(Qa,a,fma(Qb,b,S)).
__device__ __forceinline__ float filter_threshold(double d2k, float qa, float qb) {
    if (!(d2k < CMX_INF)) return CMX_INF_F;
    const double eps = 5.9604644775390625e-8;  // 2^-24
    const double rq = fabs((double)qa) + fabs((double)qb);
    const double q2 = (double)qa * (double)qa + (double)qb * (double)qb;
    const double dk = sqrt(d2k) * (1.0 + 1e-12);
    const double span = 1.4143 * dk + 2.0 * rq;
    const double delta = 1.01 * eps * span;
    const double thr = (dk + delta) * (dk + delta) - q2 + 4.04 * eps * span * span;
    float f = __double2float_ru(thr);
    f += fabsf(f) * 2.4e-7f + 1e-37f;  // strictly above, so the test can be '<'
    return f;
}

template <int QPT>
__device__ __forceinline__ bool target_coords(const KnnArgs& a, int tile, int ts, double& ta, double& tb,
                                              long long& m) {
    constexpr int TQ = WARPS * QPT * 32;
    constexpr int TR = WARPS * QPT;
    if (a.j.grid) {
        const int tr = tile / a.tiles_c;
        const int tc = tile - tr * a.tiles_c;
        const long long row = (long long)tr * TR + (ts >> 5);
        const long long col = (long long)tc * 32 + (ts & 31);
        if (row >= a.j.n_a || col >= a.j.n_b) return false;
        ta = a.j.t_a[row];
        tb = a.j.t_b[col];
        m = row * a.j.n_b + col;
    } else {
        const long long p = (long long)tile * TQ + ts;
        if (p >= a.M) return false;
        ta = a.j.t_a[p];
        tb = a.j.t_b[p];
        m = p;
    }
    return true;
}

// Warp-cooperative exact re-rank of one target's candidate buffer.
// Returns the new filter threshold (mid-sweep) or the fused IDW value (final).
template <int E>
__device__ __noinline__ double rerank_target(const KnnArgs& a, int lane, long long slot, double ta, double tb,
                                             long long m, float qa, float qb, int cnt, bool has_list,
                                             bool final) {
    constexpr int KP = 32 * E;
    double ld[E];
    int li[E];
    double* Ld = a.list_d + slot * KP;
    int* Li = a.list_i + slot * KP;
    if (has_list) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            ld[e] = Ld[e * 32 + lane];
            li[e] = Li[e * 32 + lane];
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            ld[e] = CMX_INF;
            li[e] = INT_MAX;
        }
    }
    const unsigned* C = a.cand + slot * CAP;
    for (int b = 0; b < cnt; b += 32) {
        double cd = CMX_INF;
        int ci = INT_MAX;
        if (b + lane < cnt) {
            const unsigned s = C[b + lane];
            // bit-identical to cKDTree: plain float64 dx*dx + dy*dy, no FMA contraction
            const double da = __dsub_rn(ta, a.j.s_a[s]);
            const double db = __dsub_rn(tb, a.j.s_b[s]);
            cd = __dadd_rn(__dmul_rn(da, da), __dmul_rn(db, db));
            ci = (int)s;
            if (!(cd == cd)) cd = CMX_INF;  // NaN coordinates never rank
        }
        warp_sort32(cd, ci, lane);
        merge_sorted32<E>(ld, li, cd, ci, lane);
    }
    const int k = a.j.k;
    if (!final) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            Ld[e * 32 + lane] = ld[e];
            Li[e * 32 + lane] = li[e];
        }
        double sel = ld[0];
#pragma unroll
        for (int e = 1; e < E; ++e)
            if (((k - 1) >> 5) == e) sel = ld[e];
        const double d2k = __shfl_sync(kFull, sel, (k - 1) & 31);
        return (double)filter_threshold(d2k, qa, qb);
    }

    // ---- final: neighbour table + fused IDW epilogue (idw.py:163-180) ----
    double wsum = 0.0;
    double w[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        const bool in = i < k;
        if (in) {
            if (a.j.idx_out) a.j.idx_out[m * k + i] = li[e];
            if (a.j.dist_out) a.j.dist_out[m * k + i] = sqrt(ld[e]);
            if (a.j.d2_out) a.j.d2_out[m * k + i] = ld[e];
        }
        w[e] = (in && a.j.idw_mode) ? idw_raw_weight(ld[e], a.j.power) : 0.0;
        wsum += w[e];
    }
    if (!a.j.idw_mode) return 0.0;
    wsum = warp_sum(wsum);  // weights /= nansum(weights)            idw.py:165
    double num = 0.0, den = 0.0;
    int fin = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        if (i < k) {
            const double wn = w[e] / wsum;
            const double v = a.j.value_is_f32 ? (double)static_cast<const float*>(a.j.vals)[li[e]]
                                              : static_cast<const double*>(a.j.vals)[li[e]];
            if (isfinite(v)) {  // weights[~isfinite] = 0; nansum(v*w)   idw.py:167-173
                num += v * wn;
                den += wn;
                ++fin;
            }
        }
    }
    num = warp_sum(num);
    den = warp_sum(den);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fin += __shfl_xor_sync(kFull, fin, o);
    if (fin < a.j.k_min || den == 0.0) return quiet_nan<double>();  // idw.py:179-180
    return num / den;
}

template <int E, int QPT>
__global__ void __launch_bounds__(THREADS, MAX_CTAS_PER_SM) knn_kernel(const __grid_constant__ KnnArgs a) {
    constexpr int TQ = WARPS * QPT * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* raw_a = reinterpret_cast<double*>(smem_raw);
    double* raw_b = raw_a + TS;
    float* comp = reinterpret_cast<float*>(raw_b + TS);  // [2][3][TS]
    __shared__ __align__(8) uint64_t mbar;
    __shared__ double red[4 * WARPS];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&mbar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    const long long n = a.j.n;
    const int n_st = (int)((n + TS - 1) / TS);
    const long long slot0 = (long long)blockIdx.x * TQ + (long long)(warp * QPT) * 32;  // + q*32 + lane
    unsigned* const mycand = a.cand + (slot0 + lane) * CAP;

    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        // ---------------- targets of this thread ----------------
        float Qa[QPT], Qb[QPT], thr[QPT];
        int off[QPT];
        unsigned validm = 0, has_list = 0;
        double ca, cb;
        {
            double ta[QPT], tb[QPT];
            double mn_a = CMX_INF, mx_a = -CMX_INF, mn_b = CMX_INF, mx_b = -CMX_INF;
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                long long m;
                ta[q] = 0.0;
                tb[q] = 0.0;
                if (target_coords<QPT>(a, tile, (warp * QPT + q) * 32 + lane, ta[q], tb[q], m)) {
                    validm |= 1u << q;
                    mn_a = fmin(mn_a, ta[q]);
                    mx_a = fmax(mx_a, ta[q]);
                    mn_b = fmin(mn_b, tb[q]);
                    mx_b = fmax(mx_b, tb[q]);
                }
            }
            mn_a = warp_min(mn_a);
            mx_a = warp_max(mx_a);
            mn_b = warp_min(mn_b);
            mx_b = warp_max(mx_b);
            if (lane == 0) {
                red[warp * 4 + 0] = mn_a;
                red[warp * 4 + 1] = mx_a;
                red[warp * 4 + 2] = mn_b;
                red[warp * 4 + 3] = mx_b;
            }
            __syncthreads();
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                mn_a = fmin(mn_a, red[w * 4 + 0]);
                mx_a = fmax(mx_a, red[w * 4 + 1]);
                mn_b = fmin(mn_b, red[w * 4 + 2]);
                mx_b = fmax(mx_b, red[w * 4 + 3]);
            }
            ca = 0.5 * (mn_a + mx_a);  // tile centre: keeps |q|, |s| small where it matters
            cb = 0.5 * (mn_b + mx_b);
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                const bool ok = (validm >> q) & 1u;
                Qa[q] = ok ? -2.0f * __double2float_rn(__dsub_rn(ta[q], ca)) : 0.0f;
                Qb[q] = ok ? -2.0f * __double2float_rn(__dsub_rn(tb[q], cb)) : 0.0f;
                thr[q] = ok ? CMX_INF_F : -CMX_INF_F;
                off[q] = 0;
            }
        }

        // warp-cooperative re-rank of every target of this warp selected by `pick`
        auto rerank_pass = [&](bool final, double* res) {
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                unsigned need = __ballot_sync(kFull, final ? ((validm >> q) & 1u) != 0 : off[q] > FLUSH_AT);
                while (need) {
                    const int L = __ffs(need) - 1;
                    need &= need - 1;
                    const int cnt = __shfl_sync(kFull, off[q], L);
                    const bool has = (__shfl_sync(kFull, has_list, L) >> q) & 1u;
                    const int ts = (warp * QPT + q) * 32 + L;
                    double ta = 0.0, tb = 0.0;
                    long long m = 0;
                    target_coords<QPT>(a, tile, ts, ta, tb, m);
                    const float qa = __double2float_rn(__dsub_rn(ta, ca));
                    const float qb = __double2float_rn(__dsub_rn(tb, cb));
                    const double r = rerank_target<E>(a, lane, slot0 + q * 32 + L, ta, tb, m, qa, qb, cnt, has, final);
                    if (lane == L) {
                        if (final) {
                            res[q] = r;
                        } else {
                            thr[q] = (float)r;
                            off[q] = 0;
                            has_list |= 1u << q;
                        }
                    }
                }
            }
        };

        // ---------------- sweep over all samples ----------------
        if (tid == 0 && a.aligned && n >= TS) {
            mbar_arrive_expect_tx(&mbar, 2u * TS * 8u);
            tma_load_1d(raw_a, a.j.s_a, TS * 8u, &mbar);
            tma_load_1d(raw_b, a.j.s_b, TS * 8u, &mbar);
        }
        int buf = 0;
        for (int st = 0; st < n_st; ++st) {
            const long long base = (long long)st * TS;
            const int cntS = (int)((n - base) < (long long)TS ? (n - base) : (long long)TS);
            const bool bulk = a.aligned && cntS == TS;
            float* cA = comp + buf * 3 * TS;
            float* cB = cA + TS;
            float* cS = cB + TS;
            if (bulk) {
                mbar_wait(&mbar, phase);
                phase ^= 1u;
            }
            // re-centre on the tile, round to fp32, precompute |s|^2
            for (int i = tid; i < TS; i += THREADS) {
                double xa, xb;
                if (bulk) {
                    xa = raw_a[i];
                    xb = raw_b[i];
                } else if (i < cntS) {
                    xa = a.j.s_a[base + i];
                    xb = a.j.s_b[base + i];
                } else {
                    xa = xb = quiet_nan<double>();  // padding never passes the filter
                }
                const float fa = __double2float_rn(__dsub_rn(xa, ca));
                const float fb = __double2float_rn(__dsub_rn(xb, cb));
                cA[i] = fa;
                cB[i] = fb;
                cS[i] = fmaf(fb, fb, fa * fa);
            }
            __syncthreads();  // stage complete; raw buffers free again
            if (tid == 0 && st + 1 < n_st && a.aligned && (n - (base + TS)) >= (long long)TS) {
                mbar_arrive_expect_tx(&mbar, 2u * TS * 8u);
                tma_load_1d(raw_a, a.j.s_a + base + TS, TS * 8u, &mbar);
                tma_load_1d(raw_b, a.j.s_b + base + TS, TS * 8u, &mbar);
            }

            const int cnt_pad = (cntS + CHK - 1) & ~(CHK - 1);
            const unsigned jbase = (unsigned)base;
#pragma unroll 1
            for (int c0 = 0; c0 < cnt_pad; c0 += CHK) {
#pragma unroll
                for (int j = 0; j < CHK; j += 4) {
                    const float4 A4 = *reinterpret_cast<const float4*>(cA + c0 + j);
                    const float4 B4 = *reinterpret_cast<const float4*>(cB + c0 + j);
                    const float4 S4 = *reinterpret_cast<const float4*>(cS + c0 + j);
                    const float xa[4] = {A4.x, A4.y, A4.z, A4.w};
                    const float xb[4] = {B4.x, B4.y, B4.z, B4.w};
                    const float xs[4] = {S4.x, S4.y, S4.z, S4.w};
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        float t[QPT];
                        bool any = false;
#pragma unroll
                        for (int q = 0; q < QPT; ++q) {
                            t[q] = fmaf(Qa[q], xa[s], fmaf(Qb[q], xb[s], xs[s]));
                            any |= t[q] < thr[q];
                        }
                        if (any) {
                            const unsigned jj = jbase + (unsigned)(c0 + j + s);
#pragma unroll
                            for (int q = 0; q < QPT; ++q) {
                                if (t[q] < thr[q]) {
                                    mycand[q * 32 * CAP + off[q]] = jj;
                                    ++off[q];
                                }
                            }
                        }
                    }
                }
                int mx = off[0];
#pragma unroll
                for (int q = 1; q < QPT; ++q) mx = max(mx, off[q]);
                if (__any_sync(kFull, mx > FLUSH_AT)) rerank_pass(false, nullptr);
            }
            buf ^= 1;
        }

        // ---------------- final re-rank + epilogue ----------------
        double res[QPT];
#pragma unroll
        for (int q = 0; q < QPT; ++q) res[q] = 0.0;
        rerank_pass(true, res);
        if (a.j.idw_mode == 1) {
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                if ((validm >> q) & 1u) {
                    double ta, tb;
                    long long m;
                    target_coords<QPT>(a, tile, (warp * QPT + q) * 32 + lane, ta, tb, m);
                    if (a.j.value_is_f32)
                        static_cast<float*>(a.j.out)[m] = (float)res[q];
                    else
                        static_cast<double*>(a.j.out)[m] = res[q];
                }
            }
        }
    }
}

template <int E, int QPT>
int launch(const KnnArgs& args, int ctas, cudaStream_t stream) {
    const size_t smem = 2 * TS * sizeof(double) + 6 * TS * sizeof(float);
    CMX_CUDA(cudaFuncSetAttribute(knn_kernel<E, QPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_kernel<E, QPT><<<ctas, THREADS, smem, stream>>>(args);
    note_launch();
    return check_launch();
}

inline int list_slots(int k) { return k <= 32 ? 1 : (k <= 64 ? 2 : 4); }

}  // namespace

size_t knn_workspace_bytes(int k) {
    if (k < 1 || k > CMX_MAX_K) return 0;
    const size_t ctas = (size_t)sm_count() * MAX_CTAS_PER_SM;
    const size_t kp = 32 * (size_t)list_slots(k);
    return ctas * MAX_TQ * (CAP * sizeof(unsigned) + kp * (sizeof(double) + sizeof(int))) + 256;
}

int run_knn(const KnnJob& job, cudaStream_t stream) {
    if (!job.s_a || !job.s_b || !job.t_a || !job.t_b) return CMX_ERR_BAD_ARG;
    if (job.n < 0 || job.n_a < 0 || job.n_b < 0 || job.k < 1) return CMX_ERR_BAD_ARG;
    if (!job.grid && job.n_a != job.n_b) return CMX_ERR_BAD_ARG;
    if (job.k > CMX_MAX_K) return CMX_ERR_K_TOO_LARGE;
    if ((long long)job.k > job.n) return CMX_ERR_K_GT_N;
    if (job.idw_mode == 1 && (!job.vals || !job.out || job.k_min < 1)) return CMX_ERR_BAD_ARG;
    const long long M = job.grid ? job.n_a * job.n_b : job.n_a;
    if (M == 0) return CMX_OK;
    if (job.n > 0x7fffffffLL || M > 0x7fffffffLL * 16LL) return CMX_ERR_TOO_MANY;
    const size_t need = knn_workspace_bytes(job.k);
    if (!job.workspace || job.workspace_bytes < need) return CMX_ERR_WORKSPACE;

    KnnArgs args;
    args.j = job;
    args.M = M;
    args.aligned = ((reinterpret_cast<uintptr_t>(job.s_a) | reinterpret_cast<uintptr_t>(job.s_b)) & 15u) == 0;
    const int sms = sm_count();
    const int slots = sms * MAX_CTAS_PER_SM;

    // Large tiles (8 targets per thread) amortise the shared-memory loads best; fall back to
    // 2 per thread when that would leave SMs idle.
    auto tiles_for = [&](int qpt, int& tiles_c) -> long long {
        const int tr = WARPS * qpt, tq = WARPS * qpt * 32;
        if (job.grid) {
            tiles_c = (int)((job.n_b + 31) / 32);
            return ((job.n_a + tr - 1) / tr) * (long long)tiles_c;
        }
        tiles_c = 1;
        return (M + tq - 1) / tq;
    };
    int tiles_c8 = 1, tiles_c2 = 1;
    const long long t8 = tiles_for(8, tiles_c8);
    const long long t2 = tiles_for(2, tiles_c2);
    const bool big = t8 >= 2LL * slots;
    const long long tiles = big ? t8 : t2;
    if (tiles > 0x7fffffffLL) return CMX_ERR_TOO_MANY;
    args.tiles_c = big ? tiles_c8 : tiles_c2;
    args.num_tiles = (int)tiles;
    const int ctas = (int)(tiles < slots ? tiles : slots);

    const size_t kp = 32 * (size_t)list_slots(job.k);
    unsigned char* ws = static_cast<unsigned char*>(job.workspace);
    ws = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
    const size_t nslots = (size_t)slots * MAX_TQ;
    args.list_d = reinterpret_cast<double*>(ws);
    args.cand = reinterpret_cast<unsigned*>(ws + nslots * kp * sizeof(double));
    args.list_i = reinterpret_cast<int*>(ws + nslots * kp * sizeof(double) + nslots * CAP * sizeof(unsigned));

    const int e = list_slots(job.k);
    if (big) {
        if (e == 1) return launch<1, 8>(args, ctas, stream);
        if (e == 2) return launch<2, 8>(args, ctas, stream);
        return launch<4, 8>(args, ctas, stream);
    }
    if (e == 1) return launch<1, 2>(args, ctas, stream);
    if (e == 2) return launch<2, 2>(args, ctas, stream);
    return launch<4, 2>(args, ctas, stream);
}

}  // namespace cmx
