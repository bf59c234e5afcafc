// a12 Gram-Schmidt (complex, with the reference's conjugated coefficient; real variant for inpainting) and
// a14 NPPC projection / second-moment loss statistics, as TWO HBM passes:
//   pass 1  gram_kernel : one read of the n (+1: err = gt - pred) vectors -> Hermitian Gram matrix, fp32 per-thread
//                         partials over short runs, fp64 block reduction + fp64 atomics  (fp64 Gram)
//   solve   gs_solve    : per sample, replay the reference's MGS recurrences symbolically on coefficient vectors
//                         (w_i = sum_k a_i[k] x_k) in fp64 -> lower-triangular A, norms, projections, loss terms
//   pass 2  apply_kernel: out_i = sum_{k<=i} A[i][k] x_k  (fp32 streaming FMA), one read + one write
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int NV_MAX = 13;       // up to 12 directions + err
constexpr int RUN = 16;          // elements per thread accumulated in fp32 before the fp64 reduction

struct SampleScratch {
    double G[NV_MAX * NV_MAX * 2];   // Hermitian Gram, row-major complex (re,im); real variant uses .re only
    float A[12 * 12 * 2];            // coefficient matrix (complex float), row i = direction i
};

// vector v of sample b: v < n -> x[b][v] ([2][P] for complex, [P] for real); v == n -> err = gt - pred
template <bool COMPLEX>
__device__ __forceinline__ void load_vec(const float* __restrict__ x, const float* __restrict__ gt,
                                         const float* __restrict__ pred, int b, int n, int v, long long P,
                                         long long p, float& re, float& im) {
    if (v < n) {
        const float* base = x + ((size_t)b * n + v) * (COMPLEX ? 2 : 1) * P;
        re = base[p];
        im = COMPLEX ? base[P + p] : 0.f;
    } else {
        const float* g = gt + (size_t)b * (COMPLEX ? 2 : 1) * P;
        const float* q = pred + (size_t)b * (COMPLEX ? 2 : 1) * P;
        re = g[p] - q[p];
        im = COMPLEX ? (g[P + p] - q[P + p]) : 0.f;
    }
}

// Gram rows J0..J1-1 (upper triangle, k >= j) of NV vectors. grid (chunks, B).
template <bool COMPLEX, int NV, int J0, int J1>
__device__ __forceinline__ void gram_tile(const float* __restrict__ x, const float* __restrict__ gt,
                                          const float* __restrict__ pred, int n, long long P,
                                          SampleScratch* __restrict__ scr, int b, int tile) {
    constexpr int NPAIR = (J1 - J0) * NV - (J1 * (J1 - 1) / 2 - J0 * (J0 - 1) / 2);
    constexpr int NACC = NPAIR * (COMPLEX ? 2 : 1);
    // one tile = one run of TPB*RUN elements: fp32 per-thread partials over RUN (=16) products only, fp64 from there on
    const long long base = (long long)tile * TPB * RUN;
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int r = 0; r < RUN; ++r) {
        long long p = base + (long long)r * TPB + threadIdx.x;
        if (p < P) {
            float vr[NV], vi[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) load_vec<COMPLEX>(x, gt, pred, b, n, v, P, p, vr[v], vi[v]);
            int a = 0;
#pragma unroll
            for (int j = J0; j < J1; ++j) {
#pragma unroll
                for (int k = j; k < NV; ++k) {
                    if (COMPLEX) {
                        acc[a] += vr[j] * vr[k] + vi[j] * vi[k];      // Re conj(x_j) x_k
                        acc[a + 1] += vr[j] * vi[k] - vi[j] * vr[k];  // Im conj(x_j) x_k
                        a += 2;
                    } else {
                        acc[a] += vr[j] * vr[k];
                        a += 1;
                    }
                }
            }
        }
    }
    // block reduction: fp32 shuffles inside a warp (32 partials of <=16 products each), fp64 across the 8 warps,
    // one fp64 atomic per Gram entry per CTA
    __shared__ float wsum[TPB / 32][NACC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float v = nppc::warp_sum(acc[i]);
        if (lane == 0) wsum[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        double d = 0.0;
#pragma unroll
        for (int w = 0; w < TPB / 32; ++w) d += (double)wsum[w][threadIdx.x];
        // accumulator index -> (j, k, re/im)
        int a = threadIdx.x, comp = COMPLEX ? (a & 1) : 0, pair = COMPLEX ? (a >> 1) : a;
        int j = J0, rowlen = NV - J0;
        while (pair >= rowlen) { pair -= rowlen; ++j; --rowlen; }
        int k = j + pair;
        atomicAdd(&scr[b].G[(j * NV_MAX + k) * 2 + comp], d);
    }
}
template <bool COMPLEX, int NV, int J0, int J1>
__global__ void __launch_bounds__(TPB) gram_kernel(const float* __restrict__ x, const float* __restrict__ gt,
                                                  const float* __restrict__ pred, int n, long long P,
                                                  SampleScratch* __restrict__ scr) {
    gram_tile<COMPLEX, NV, J0, J1>(x, gt, pred, n, P, scr, blockIdx.y, blockIdx.x);
}

struct cd { double x, y; };
__device__ __forceinline__ cd cmul(cd a, cd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd cconj(cd a) { return {a.x, -a.y}; }

// One thread per sample. mode: do_gs (orthogonalise; else A = identity), has_err (loss statistics).
template <bool COMPLEX, int N_T>   // N_T = compile-time n (fully unrolled, everything in registers); 0 = generic
__global__ void gs_solve_kernel(SampleScratch* __restrict__ scr, int B, int n_rt, int do_gs, int has_err,
                                float* __restrict__ err_norm, float* __restrict__ err_proj,
                                float* __restrict__ w_norms, float* __restrict__ reconst_err,
                                float* __restrict__ second_moment) {
    int b = blockIdx.x;
    if (b >= B || threadIdx.x != 0) return;
    SampleScratch& s = scr[b];
    const int n = N_T > 0 ? N_T : n_rt;
    const int nv = n + (has_err ? 1 : 0);
    constexpr int GD = N_T > 0 ? N_T + 1 : NV_MAX;
    cd Gl[GD][GD];   // full Hermitian matrix from the stored upper triangle (local copy: no global loads in the solve)
    for (int j = 0; j < nv; ++j)
        for (int k = j; k < nv; ++k) {
            cd g{s.G[(j * NV_MAX + k) * 2], s.G[(j * NV_MAX + k) * 2 + 1]};
            Gl[j][k] = g;
            Gl[k][j] = cd{g.x, -g.y};
        }
    auto G = [&](int j, int k) -> cd { return Gl[j][k]; };
    constexpr int ND = N_T > 0 ? N_T : 12;
    cd ahat[ND][ND];  // normalised coefficient vectors of the previous directions
    cd v[ND][ND];     // v_j = G * ahat_j
    double nu[ND];
    for (int i = 0; i < n; ++i) {
        cd a[ND];
        for (int k = 0; k < n; ++k) a[k] = cd{k == i ? 1.0 : 0.0, 0.0};
        if (do_gs) {
            for (int j = 0; j < i; ++j) {
                // reference coefficient: c = sum_p conj(w[p]) * what_j[p] = sum_k conj(a[k]) (G ahat_j)[k]
                cd c{0.0, 0.0};
                for (int k = 0; k <= i; ++k) { cd t = cmul(cconj(a[k]), v[j][k]); c.x += t.x; c.y += t.y; }
                for (int k = 0; k <= j; ++k) { cd t = cmul(ahat[j][k], c); a[k].x -= t.x; a[k].y -= t.y; }
            }
        }
        // ||w_i||^2 = a^H G a
        double nrm2 = 0.0;
        for (int k = 0; k <= i; ++k) {
            cd gk{0.0, 0.0};
            for (int l = 0; l <= i; ++l) { cd t = cmul(G(k, l), a[l]); gk.x += t.x; gk.y += t.y; }
            cd t = cmul(cconj(a[k]), gk);
            nrm2 += t.x;
        }
        double nrm = sqrt(nrm2 > 0.0 ? nrm2 : 0.0);
        nu[i] = nrm;
        for (int k = 0; k < n; ++k) {
            ahat[i][k] = (k <= i) ? cd{a[k].x / nrm, a[k].y / nrm} : cd{0.0, 0.0};  // no epsilon (pc_wrapper.py:37)
            s.A[(i * 12 + k) * 2] = (k <= i) ? (float)a[k].x : 0.f;
            s.A[(i * 12 + k) * 2 + 1] = (k <= i) ? (float)a[k].y : 0.f;
        }
        for (int k = 0; k < n; ++k) {
            cd gk{0.0, 0.0};
            for (int l = 0; l <= i; ++l) { cd t = cmul(G(k, l), ahat[i][l]); gk.x += t.x; gk.y += t.y; }
            v[i][k] = gk;
        }
        if (has_err) {
            // complex: trainer.py:270-295 (nu/(eps+1e-8), W/(nu+1e-8));  real: inpainting nppc_trainer.py:354-373
            // (1e-6 is added to BOTH norms before any division, and the logged err_norm includes it)
            double eps_n = sqrt(fmax(G(n, n).x, 0.0));
            cd pr{0.0, 0.0};
            for (int k = 0; k <= i; ++k) { cd t = cmul(cconj(a[k]), G(k, n)); pr.x += t.x; pr.y += t.y; }
            const double den = COMPLEX ? (nrm + 1e-8) * (eps_n + 1e-8) : (nrm + 1e-6) * (eps_n + 1e-6);
            pr.x /= den; pr.y /= den;
            if (COMPLEX) {
                err_proj[((size_t)b * n + i) * 2] = (float)pr.x;
                err_proj[((size_t)b * n + i) * 2 + 1] = (float)pr.y;
            } else {
                err_proj[(size_t)b * n + i] = (float)pr.x;
            }
            double wn = COMPLEX ? nrm / (eps_n + 1e-8) : (nrm + 1e-6) / (eps_n + 1e-6);
            w_norms[(size_t)b * n + i] = (float)wn;
            double pm2 = pr.x * pr.x + pr.y * pr.y;
            double d = wn * wn - pm2;
            second_moment[(size_t)b * n + i] = (float)(d * d);
            nu[i] = pm2;
        }
    }
    if (has_err) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += nu[i];
        reconst_err[b] = (float)(1.0 - acc);
        err_norm[b] = (float)(sqrt(fmax(G(n, n).x, 0.0)) + (COMPLEX ? 0.0 : 1e-6));
    }
}

// Warp-parallel version of the solve: one warp per sample, lane k owns coefficient k.  Same recurrences as above, but
// the O(n) inner sums are warp reductions / shuffles, so the serial fp64 dependency chain shrinks from O(n^3) to
// O(n^2 log 32) steps (n = 5: 28 us -> ~5 us; the solve sits between the two HBM passes, so its latency is exposed).
__device__ __forceinline__ cd warp_sum_cd(cd v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}
struct SolveSmem { cd Gs[NV_MAX][NV_MAX + 1], ahat_s[12][NV_MAX + 1], v_s[12][NV_MAX + 1]; };
// one warp solves sample b (lane k owns coefficient k)
template <bool COMPLEX>
__device__ __noinline__ void solve_warp(SampleScratch* __restrict__ scr, int b, int n, int do_gs, int has_err,
                                           float* __restrict__ err_norm, float* __restrict__ err_proj,
                                           float* __restrict__ w_norms, float* __restrict__ reconst_err,
                                           float* __restrict__ second_moment, SolveSmem& sm, int lane) {
    auto& Gs = sm.Gs;
    auto& ahat_s = sm.ahat_s;
    auto& v_s = sm.v_s;
    SampleScratch& s = scr[b];
    const int nv = n + (has_err ? 1 : 0);
    for (int idx = lane; idx < nv * nv; idx += 32) {
        const int j = idx / nv, k = idx - j * nv;
        Gs[j][k] = (k >= j) ? cd{s.G[(j * NV_MAX + k) * 2], s.G[(j * NV_MAX + k) * 2 + 1]}
                            : cd{s.G[(k * NV_MAX + j) * 2], -s.G[(k * NV_MAX + j) * 2 + 1]};
    }
    __syncwarp();
    const cd zero{0.0, 0.0};
    const double eps_n = has_err ? sqrt(fmax(Gs[n][n].x, 0.0)) : 0.0;
    double nu_sum = 0.0;
    for (int i = 0; i < n; ++i) {
        cd a = (lane == i) ? cd{1.0, 0.0} : zero;   // lane k holds a[k]
        if (do_gs) {
            for (int j = 0; j < i; ++j) {
                // reference coefficient: c = sum_p conj(w[p]) * what_j[p] = sum_k conj(a[k]) (G ahat_j)[k]
                const cd c = warp_sum_cd(lane <= i ? cmul(cconj(a), v_s[j][lane]) : zero);
                if (lane <= j) { const cd t = cmul(ahat_s[j][lane], c); a.x -= t.x; a.y -= t.y; }
            }
        }
        // gk = (G a)[lane] for every row (also the err row n); a[l] is broadcast from lane l
        cd gk = zero;
        for (int l = 0; l <= i; ++l) {
            const cd al{__shfl_sync(0xffffffffu, a.x, l), __shfl_sync(0xffffffffu, a.y, l)};
            if (lane < nv) { const cd t = cmul(Gs[lane][l], al); gk.x += t.x; gk.y += t.y; }
        }
        const cd q = warp_sum_cd(lane <= i ? cmul(cconj(a), gk) : zero);   // ||w_i||^2 = a^H G a
        const double nrm = sqrt(q.x > 0.0 ? q.x : 0.0);
        if (lane < 12) {
            ahat_s[i][lane] = (lane <= i) ? cd{a.x / nrm, a.y / nrm} : zero;   // no epsilon (pc_wrapper.py:37)
            v_s[i][lane] = (lane < n) ? cd{gk.x / nrm, gk.y / nrm} : zero;     // v_i = G ahat_i
        }
        if (lane < n) {
            s.A[(i * 12 + lane) * 2] = (lane <= i) ? (float)a.x : 0.f;
            s.A[(i * 12 + lane) * 2 + 1] = (lane <= i) ? (float)a.y : 0.f;
        }
        if (has_err) {
            // complex: trainer.py:270-295 (nu/(eps+1e-8), W/(nu+1e-8));  real: inpainting nppc_trainer.py:354-373
            cd pr = warp_sum_cd(lane <= i ? cmul(cconj(a), Gs[lane][n]) : zero);
            const double den = COMPLEX ? (nrm + 1e-8) * (eps_n + 1e-8) : (nrm + 1e-6) * (eps_n + 1e-6);
            pr.x /= den; pr.y /= den;
            const double wn = COMPLEX ? nrm / (eps_n + 1e-8) : (nrm + 1e-6) / (eps_n + 1e-6);
            const double pm2 = pr.x * pr.x + pr.y * pr.y;
            const double d = wn * wn - pm2;
            nu_sum += pm2;
            if (lane == 0) {
                if (COMPLEX) {
                    err_proj[((size_t)b * n + i) * 2] = (float)pr.x;
                    err_proj[((size_t)b * n + i) * 2 + 1] = (float)pr.y;
                } else {
                    err_proj[(size_t)b * n + i] = (float)pr.x;
                }
                w_norms[(size_t)b * n + i] = (float)wn;
                second_moment[(size_t)b * n + i] = (float)(d * d);
            }
        }
        __syncwarp();
    }
    if (has_err && lane == 0) {
        reconst_err[b] = (float)(1.0 - nu_sum);
        err_norm[b] = (float)(eps_n + (COMPLEX ? 0.0 : 1e-6));
    }
}
template <bool COMPLEX>
__global__ void __launch_bounds__(32) gs_solve_warp_kernel(SampleScratch* __restrict__ scr, int B, int n, int do_gs, int has_err,
                                                          float* __restrict__ err_norm, float* __restrict__ err_proj,
                                                          float* __restrict__ w_norms, float* __restrict__ reconst_err,
                                                          float* __restrict__ second_moment) {
    __shared__ SolveSmem sm;
    if ((int)blockIdx.x >= B) return;
    solve_warp<COMPLEX>(scr, blockIdx.x, n, do_gs, has_err, err_norm, err_proj, w_norms, reconst_err, second_moment, sm, threadIdx.x);
}

// ---------------------------------------------------------------------------------------------------------------------
// Single-launch pipeline: Gram pass, coefficient solve and apply pass of ALL samples in ONE persistent kernel.
// Tasks are handed out in order by an atomic counter: stage s = Gram tiles of sample s, then apply tiles of sample s - LAG.
// The CTA that finishes the last Gram tile of a sample solves it (one warp, fp64) and publishes `ready[b]`; by the time the
// apply tiles of that sample come up (LAG samples of work later) the solve is long done and the sample's 2.6 MB are still in
// the 126 MB L2, so every vector is read from HBM once and written once (the three-launch version re-read x from HBM for
// the apply pass: 495 MB moved for 330 MB algorithmic at B = 64) and the solve's latency hides behind other samples' tiles.
// All CTAs are co-resident (grid <= SMs x occupancy) and a task only waits on tasks handed out before it: no deadlock.
// LAG: with ~150 tasks in flight (one per CTA) and 16 tiles per sample and phase, tasks handed out together span ~5 stages; the
// apply tiles of a sample must come up well after its solve (~10 us on one warp) has finished, or whole waves of CTAs spin on
// `ready` (measured: LAG = 2 -> 395 us, slower than three launches).  8 samples = 21 MB of x, far inside the 126 MB L2.
constexpr int PIPE_LAG = 8;
struct PipeCtl { unsigned int next, pad[3]; };
template <bool COMPLEX, int NV>
__global__ void __launch_bounds__(TPB, 3) gs_pipeline_kernel(const float* __restrict__ x, const float* __restrict__ gt,
                                                         const float* __restrict__ pred, int B, int n, long long P,
                                                         SampleScratch* __restrict__ scr, PipeCtl* __restrict__ ctl,
                                                         unsigned int* __restrict__ done, unsigned int* __restrict__ ready,
                                                         int do_gs, int has_err, float* __restrict__ out,
                                                         float* __restrict__ err_norm, float* __restrict__ err_proj,
                                                         float* __restrict__ w_norms, float* __restrict__ reconst_err,
                                                         float* __restrict__ second_moment) {
    constexpr int N = 12;   // coefficient matrix stride
    __shared__ SolveSmem sm;
    __shared__ float2 A[N][N];
    __shared__ unsigned int s_task, s_last;
    const int tiles = (int)((P + (long long)TPB * RUN - 1) / ((long long)TPB * RUN));
    const int LAG = B < PIPE_LAG ? B : PIPE_LAG;
    const long long total = 2LL * B * tiles;
    const size_t vs = (size_t)(COMPLEX ? 2 : 1) * P;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_task = atomicAdd(&ctl->next, 1u);
        __syncthreads();
        const long long t = s_task;
        if (t >= total) break;
        // decode: stages 0..LAG-1 Gram only, LAG..B-1 Gram(s) then apply(s - LAG), B..B+LAG-1 apply only
        int b, tile;
        bool is_gram;
        const long long head = (long long)LAG * tiles, mid = (long long)(B - LAG) * 2 * tiles;
        if (t < head) { b = (int)(t / tiles); tile = (int)(t % tiles); is_gram = true; }
        else if (t < head + mid) {
            const long long u = t - head;
            const int sidx = LAG + (int)(u / (2 * tiles)), r = (int)(u % (2 * tiles));
            is_gram = r < tiles;
            b = is_gram ? sidx : sidx - LAG;
            tile = is_gram ? r : r - tiles;
        } else { const long long u = t - head - mid; b = B - LAG + (int)(u / tiles); tile = (int)(u % tiles); is_gram = false; }
        if (is_gram) {
            gram_tile<COMPLEX, NV, 0, NV>(x, gt, pred, n, P, scr, b, tile);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) s_last = (atomicAdd(&done[b], 1u) == (unsigned)(tiles - 1));
            __syncthreads();
            if (s_last) {
                __threadfence();
                if (threadIdx.x < 32) {
                    solve_warp<COMPLEX>(scr, b, n, do_gs, has_err, err_norm, err_proj, w_norms, reconst_err, second_moment, sm, threadIdx.x);
                    __threadfence();
                    __syncwarp();
                    if (threadIdx.x == 0) atomicExch(&ready[b], 1u);
                }
            }
        } else {
            if (threadIdx.x == 0) {
                while (atomicAdd(&ready[b], 0u) == 0u) __nanosleep(64);
                __threadfence();
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
                const int r = i / n, c = i % n;
                A[r][c] = make_float2(__ldcg(&scr[b].A[(r * 12 + c) * 2]), __ldcg(&scr[b].A[(r * 12 + c) * 2 + 1]));
            }
            __syncthreads();
            const float* xb = x + (size_t)b * n * vs;
            float* ob = out + (size_t)b * n * vs;
            const long long p0 = (long long)tile * TPB * RUN;
#pragma unroll 4
            for (int r = 0; r < RUN; ++r) {
                const long long p = p0 + (long long)r * TPB + threadIdx.x;
                if (p >= P) break;
                float xr[NV], xi[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    if (k < n) { xr[k] = xb[k * vs + p]; xi[k] = COMPLEX ? xb[k * vs + P + p] : 0.f; }
                }
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    if (i >= n) break;
                    float wr = 0.f, wi = 0.f;
#pragma unroll
                    for (int k = 0; k <= i; ++k) {
                        const float2 a = A[i][k];
                        if (COMPLEX) { wr += a.x * xr[k] - a.y * xi[k]; wi += a.x * xi[k] + a.y * xr[k]; }
                        else wr += a.x * xr[k];
                    }
                    if (i == 0) { wr = xr[0]; wi = xi[0]; }   // direction 0 is returned untouched (bit-exact)
                    __stcs(&ob[i * vs + p], wr);
                    if (COMPLEX) __stcs(&ob[i * vs + P + p], wi);
                }
            }
        }
    }
}

// out_i[p] = sum_{k<=i} A[i][k] x_k[p]. grid (chunks, B)
template <bool COMPLEX, int N>
__global__ void __launch_bounds__(TPB) apply_kernel(const float* __restrict__ x, long long P,
                                                   const SampleScratch* __restrict__ scr, float* __restrict__ out) {
    __shared__ float2 A[N][N];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
        int r = i / N, c = i % N;
        A[r][c] = make_float2(scr[b].A[(r * 12 + c) * 2], scr[b].A[(r * 12 + c) * 2 + 1]);
    }
    __syncthreads();
    const size_t vs = (size_t)(COMPLEX ? 2 : 1) * P;
    const float* xb = x + (size_t)b * N * vs;
    float* ob = out + (size_t)b * N * vs;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        float xr[N], xi[N];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            xr[k] = xb[k * vs + p];
            xi[k] = COMPLEX ? xb[k * vs + P + p] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float wr = 0.f, wi = 0.f;
#pragma unroll
            for (int k = 0; k <= i; ++k) {
                float2 a = A[i][k];
                if (COMPLEX) {
                    wr += a.x * xr[k] - a.y * xi[k];
                    wi += a.x * xi[k] + a.y * xr[k];
                } else {
                    wr += a.x * xr[k];
                }
            }
            if (i == 0) { wr = xr[0]; wi = xi[0]; }  // direction 0 is returned untouched (bit-exact)
            ob[i * vs + p] = wr;
            if (COMPLEX) ob[i * vs + P + p] = wi;
        }
    }
}

int chunks_for(long long P, int B) {
    (void)B;
    long long per = ((long long)TPB * RUN);
    long long c = (P + per - 1) / per;   // one CTA per run of TPB*RUN elements
    return (int)(c < 1 ? 1 : c);
}

template <bool COMPLEX, int NV>
int launch_gram(const float* x, const float* gt, const float* pred, int B, int n, long long P, SampleScratch* scr,
                cudaStream_t s) {
    dim3 grid(chunks_for(P, B), B);
    constexpr int NPAIR = NV * (NV + 1) / 2;
    if constexpr (NPAIR * (COMPLEX ? 2 : 1) <= 72) {
        gram_kernel<COMPLEX, NV, 0, NV><<<grid, TPB, 0, s>>>(x, gt, pred, n, P, scr);
        NPPC_COUNT_LAUNCH(1);
    } else {  // split the upper triangle in two row bands to stay in registers
        constexpr int JM = NV / 3 > 0 ? NV / 3 : 1;
        gram_kernel<COMPLEX, NV, 0, JM><<<grid, TPB, 0, s>>>(x, gt, pred, n, P, scr);
        gram_kernel<COMPLEX, NV, JM, NV><<<grid, TPB, 0, s>>>(x, gt, pred, n, P, scr);
        NPPC_COUNT_LAUNCH(2);
    }
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

template <bool COMPLEX>
int dispatch_gram(int NV, const float* x, const float* gt, const float* pred, int B, int n, long long P,
                  SampleScratch* scr, cudaStream_t s) {
    switch (NV) {
#define C(v) case v: return launch_gram<COMPLEX, v>(x, gt, pred, B, n, P, scr, s);
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13)
#undef C
    }
    nppc::set_error("gram: unsupported vector count %d", NV);
    return NPPC_ERR_UNSUPPORTED;
}

template <bool COMPLEX>
int dispatch_apply(int n, const float* x, int B, long long P, const SampleScratch* scr, float* out, cudaStream_t s) {
    long long c = (P + TPB - 1) / TPB;
    long long want = (16LL * nppc::sm_count() + B - 1) / B;
    if (c > want) c = want;
    dim3 grid((unsigned)(c < 1 ? 1 : c), B);
    switch (n) {
#define C(v) case v: apply_kernel<COMPLEX, v><<<grid, TPB, 0, s>>>(x, P, scr, out); break;
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12)
#undef C
        default:
            nppc::set_error("gram_schmidt: unsupported n_dirs %d (1..12)", n);
            return NPPC_ERR_UNSUPPORTED;
    }
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

template <bool COMPLEX>
int run_chunk(const float* x, const float* gt, const float* pred, int B, int n, long long P, SampleScratch* scr, int do_gs,
              float* out, float* err_norm, float* err_proj, float* w_norms, float* reconst_err, float* second_moment,
              cudaStream_t s) {
    const int has_err = gt != nullptr;
    int rc = dispatch_gram<COMPLEX>(n + has_err, x, gt, pred, B, n, P, scr, s);
    if (rc) return rc;
    static const bool serial_solve = getenv("NPPC_GS_SERIAL_SOLVE") && atoi(getenv("NPPC_GS_SERIAL_SOLVE")) != 0;
    if (serial_solve)   // the single-thread reference implementation of the same recurrences (debugging aid)
        gs_solve_kernel<COMPLEX, 0><<<B, 32, 0, s>>>(scr, B, n, do_gs, has_err, err_norm, err_proj, w_norms, reconst_err, second_moment);
    else
        gs_solve_warp_kernel<COMPLEX><<<B, 32, 0, s>>>(scr, B, n, do_gs, has_err, err_norm, err_proj, w_norms, reconst_err, second_moment);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    if (out) return dispatch_apply<COMPLEX>(n, x, B, P, scr, out, s);
    return NPPC_OK;
}

// Optional L2-resident chunking (NPPC_GS_L2_MB=<MiB>): walk the batch in chunks of samples whose vectors fit in L2 so the
// apply pass re-reads from L2.  Measured on B200 (B = 64, n = 5): 146 us unchunked vs 161 / 226 / 382 us at 80 / 40 / 20 MiB —
// every chunk pays the serial fp64 coefficient solve (~28 us) and a small-grid launch, so the default is ONE chunk.
template <bool COMPLEX, int NV>
int launch_pipeline(const float* x, const float* gt, const float* pred, int B, int n, long long P, SampleScratch* scr, int do_gs,
                    float* out, float* err_norm, float* err_proj, float* w_norms, float* reconst_err, float* second_moment,
                    cudaStream_t s) {
    // control block lives behind the per-sample scratch (nppc_gs_scratch_bytes reserves it): counter, done[B], ready[B]
    PipeCtl* ctl = reinterpret_cast<PipeCtl*>(scr + B);
    unsigned int* done = reinterpret_cast<unsigned int*>(ctl + 1);
    unsigned int* ready = done + B;
    NPPC_CUDA_OK(cudaMemsetAsync(ctl, 0, sizeof(PipeCtl) + sizeof(unsigned int) * 2 * (size_t)B, s));
    auto kern = gs_pipeline_kernel<COMPLEX, NV>;
    static int occ = 0;
    if (occ == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TPB, 0) != cudaSuccess || occ < 1) occ = 1;
    }
    const long long total = 2LL * B * chunks_for(P, B);
    long long grid = (long long)nppc::sm_count() * occ;
    if (grid > total) grid = total;
    kern<<<(unsigned)grid, TPB, 0, s>>>(x, gt, pred, B, n, P, scr, ctl, done, ready, do_gs, gt != nullptr, out, err_norm, err_proj,
                                        w_norms, reconst_err, second_moment);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

template <bool COMPLEX>
int run(const float* x, const float* gt, const float* pred, int B, int n, long long P, void* scratch, int do_gs,
        float* out, float* err_norm, float* err_proj, float* w_norms, float* reconst_err, float* second_moment,
        cudaStream_t s) {
    SampleScratch* scr = (SampleScratch*)scratch;
    NPPC_CUDA_OK(cudaMemsetAsync(scr, 0, sizeof(SampleScratch) * (size_t)B, s));
    // NPPC_GS_PIPELINE=1 opts into the single persistent launch.  It is NOT the default: measured on B200 at B = 64, n = 5
    // (tools/kernel_bench.py, profiles/r02_hbm_ncu_summary.csv) Gram-Schmidt alone takes 166 us pipelined vs 142 us in three
    // launches and 192-223 us vs 208 us with the error vector.  Two reasons, both visible in ncu: (1) the Gram tiles are
    // latency-bound (~30 us per 4096-element tile incl. the fp64 block reduction + 42 fp64 atomics) and need the 4 CTAs/SM the
    // small kernels reach — the merged kernel gets 3 (80 registers); (2) the apply pass does NOT find x in L2 (hit rate 27 %,
    // 373 MB read for 198 MB algorithmic) although only ~50 MB of other traffic separates the two passes, with or without
    // L2::evict_last / evict_first cache hints on the loads (tried: 172 / 223 us).
    static const int pipe_env = getenv("NPPC_GS_PIPELINE") ? atoi(getenv("NPPC_GS_PIPELINE")) : 0;
    const int NVr = n + (gt != nullptr ? 1 : 0);
    const bool pipeline = pipe_env == 1;
    if (pipeline && out && NVr <= 7) {   // one persistent launch (register budget of the Gram tile: up to 7 vectors)
        switch (NVr) {
#define C(v) case v: return launch_pipeline<COMPLEX, v>(x, gt, pred, B, n, P, scr, do_gs, out, err_norm, err_proj, w_norms, reconst_err, second_moment, s);
            C(1) C(2) C(3) C(4) C(5) C(6) C(7)
#undef C
        }
    }
    static const long long budget = getenv("NPPC_GS_L2_MB") ? atoll(getenv("NPPC_GS_L2_MB")) << 20 : (1LL << 50);
    const long long per_sample = (long long)n * (COMPLEX ? 2 : 1) * P * 4;
    int chunk = out ? (int)(budget / (per_sample > 0 ? per_sample : 1)) : B;   // no apply pass -> nothing to re-read
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
    const size_t vs = (size_t)(COMPLEX ? 2 : 1) * P;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int bc = B - b0 < chunk ? B - b0 : chunk;
        int rc = run_chunk<COMPLEX>(x + (size_t)b0 * n * vs, gt ? gt + (size_t)b0 * vs : nullptr, pred ? pred + (size_t)b0 * vs : nullptr,
                                    bc, n, P, scr + b0, do_gs, out ? out + (size_t)b0 * n * vs : nullptr,
                                    err_norm ? err_norm + b0 : nullptr, err_proj ? err_proj + (size_t)b0 * n * (COMPLEX ? 2 : 1) : nullptr,
                                    w_norms ? w_norms + (size_t)b0 * n : nullptr, reconst_err ? reconst_err + b0 : nullptr,
                                    second_moment ? second_moment + (size_t)b0 * n : nullptr, s);
        if (rc) return rc;
    }
    return NPPC_OK;
}

}  // namespace

// out_i = sum_{k<n} C[i][k] x_k + C[i][n] (gt - pred), complex coefficients: the one streaming pass of the Gram-Schmidt + loss
// BACKWARD (gs_backward.py: the gradient w.r.t. every head output lies in the span of the n directions and the error vector).
// grid (chunks, B); coefficients [B][n][n+1][2] fp32 staged in shared memory.
__global__ void __launch_bounds__(TPB) complex_lincomb_kernel(const float* __restrict__ x, const float* __restrict__ gt,
                                                             const float* __restrict__ pred, int n, long long P,
                                                             const float* __restrict__ coef, float* __restrict__ out) {
    __shared__ float2 C[12 * 13];
    const int b = blockIdx.y, n1 = n + 1;
    for (int i = threadIdx.x; i < n * n1; i += blockDim.x)
        C[i] = make_float2(coef[((size_t)b * n * n1 + i) * 2], coef[((size_t)b * n * n1 + i) * 2 + 1]);
    __syncthreads();
    const float* xb = x + (size_t)b * n * 2 * P;
    float* ob = out + (size_t)b * n * 2 * P;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        float vr[13], vi[13];
        for (int k = 0; k < n; ++k) { vr[k] = xb[(size_t)k * 2 * P + p]; vi[k] = xb[(size_t)k * 2 * P + P + p]; }
        load_vec<true>(x, gt, pred, b, n, n, P, p, vr[n], vi[n]);
        for (int i = 0; i < n; ++i) {
            float wr = 0.f, wi = 0.f;
            for (int k = 0; k < n1; ++k) {
                const float2 c = C[i * n1 + k];
                wr += c.x * vr[k] - c.y * vi[k];
                wi += c.x * vi[k] + c.y * vr[k];
            }
            ob[(size_t)i * 2 * P + p] = wr;
            ob[(size_t)i * 2 * P + P + p] = wi;
        }
    }
}

extern "C" int nppc_complex_lincomb(const float* x, const float* gt, const float* pred, int B, int n, long long P,
                                    const float* coef, float* out, void* stream) {
    NPPC_CHECK_ARG(x && gt && pred && coef && out && B > 0 && P > 0 && n >= 1 && n <= 12, "nppc_complex_lincomb: bad arguments");
    NPPC_CHECK_ARG(B <= 65535, "nppc_complex_lincomb: B too large");
    int chunks = nppc::cdiv((long long)nppc::sm_count() * 8, B);
    const int cap = nppc::cdiv(P, TPB);
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    complex_lincomb_kernel<<<dim3(chunks, B), TPB, 0, (cudaStream_t)stream>>>(x, gt, pred, n, P, coef, out);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" size_t nppc_gs_scratch_bytes(int B, int n) {
    (void)n;
    const size_t b = (size_t)(B > 0 ? B : 0);
    return sizeof(SampleScratch) * b + sizeof(PipeCtl) + sizeof(unsigned int) * 2 * b + 16;   // + the pipeline's control block
}

#define GS_ARGS_OK(name)                                                                          \
    NPPC_CHECK_ARG(x && scratch && B > 0 && P > 0, name ": bad arguments");                        \
    NPPC_CHECK_ARG(n >= 1 && n <= 12, name ": n_dirs must be in 1..12 (got %d)", n)

extern "C" int nppc_gram_schmidt_complex(const float* x, int B, int n, long long P, void* scratch, float* out, void* stream) {
    GS_ARGS_OK("nppc_gram_schmidt_complex");
    NPPC_CHECK_ARG(out != nullptr, "nppc_gram_schmidt_complex: null output");
    return run<true>(x, nullptr, nullptr, B, n, P, scratch, 1, out, nullptr, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int nppc_gram_schmidt_real(const float* x, int B, int n, long long P, void* scratch, float* out, void* stream) {
    GS_ARGS_OK("nppc_gram_schmidt_real");
    NPPC_CHECK_ARG(out != nullptr, "nppc_gram_schmidt_real: null output");
    return run<false>(x, nullptr, nullptr, B, n, P, scratch, 1, out, nullptr, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int nppc_gs_loss_fused(const float* x, const float* gt, const float* pred, int B, int n, long long P,
                                  void* scratch, float* w_mat, float* err_norm, float* err_proj, float* w_norms,
                                  float* reconst_err, float* second_moment_mse, void* stream) {
    GS_ARGS_OK("nppc_gs_loss_fused");
    NPPC_CHECK_ARG(gt && pred && w_mat && err_norm && err_proj && w_norms && reconst_err && second_moment_mse,
                   "nppc_gs_loss_fused: null pointer");
    return run<true>(x, gt, pred, B, n, P, scratch, 1, w_mat, err_norm, err_proj, w_norms, reconst_err, second_moment_mse,
                     (cudaStream_t)stream);
}

extern "C" int nppc_projection_loss(const float* x, const float* gt, const float* pred, int B, int n, long long P,
                                    void* scratch, float* err_norm, float* err_proj, float* w_norms,
                                    float* reconst_err, float* second_moment_mse, void* stream) {
    GS_ARGS_OK("nppc_projection_loss");
    NPPC_CHECK_ARG(gt && pred && err_norm && err_proj && w_norms && reconst_err && second_moment_mse,
                   "nppc_projection_loss: null pointer");
    return run<true>(x, gt, pred, B, n, P, scratch, 0, nullptr, err_norm, err_proj, w_norms, reconst_err, second_moment_mse,
                     (cudaStream_t)stream);
}

// Real (inpainting) variants: x [B,n,P], gt / pred [B,P]; err_proj is [B,n] real.
extern "C" int nppc_gs_loss_fused_real(const float* x, const float* gt, const float* pred, int B, int n, long long P,
                                       void* scratch, float* w_mat, float* err_norm, float* err_proj, float* w_norms,
                                       float* reconst_err, float* second_moment_mse, void* stream) {
    GS_ARGS_OK("nppc_gs_loss_fused_real");
    NPPC_CHECK_ARG(gt && pred && w_mat && err_norm && err_proj && w_norms && reconst_err && second_moment_mse,
                   "nppc_gs_loss_fused_real: null pointer");
    return run<false>(x, gt, pred, B, n, P, scratch, 1, w_mat, err_norm, err_proj, w_norms, reconst_err, second_moment_mse,
                      (cudaStream_t)stream);
}

extern "C" int nppc_projection_loss_real(const float* x, const float* gt, const float* pred, int B, int n, long long P,
                                         void* scratch, float* err_norm, float* err_proj, float* w_norms,
                                         float* reconst_err, float* second_moment_mse, void* stream) {
    GS_ARGS_OK("nppc_projection_loss_real");
    NPPC_CHECK_ARG(gt && pred && err_norm && err_proj && w_norms && reconst_err && second_moment_mse,
                   "nppc_projection_loss_real: null pointer");
    return run<false>(x, gt, pred, B, n, P, scratch, 0, nullptr, err_norm, err_proj, w_norms, reconst_err, second_moment_mse,
                      (cudaStream_t)stream);
}
