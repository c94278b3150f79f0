// engine.cu -- host side of the Fock-build engine and the C ABI of include/qcfock.h.
//
// qcf_create   : shell pairs -> (la,lb,K) groups -> Schwarz factors on the GPU -> sorted SoA pair data in HBM on every
//                device of the context, cost-balanced bra split, launch plan
//                (replaces the one-off molint::eri call, core/src/hf/rhf.rs:45, uhf.rs:55)
// qcf_build_*  : density scaling + shell-block maxima -> one eri_jk launch per (bra group, ket group), replayed as ONE
//                CUDA graph per device -> symmetrise / combine, summing the partial matrices of all devices over
//                NVLink peer memory inside the finalize kernel  (replaces rhf.rs:58-62,152-167 and uhf.rs:210-227)
// A build never waits for the device on the host: the global density maximum that steers the ket cut-offs and the
// fixed-point scale of the deterministic mode stay in HBM (BuildScalars).
// There is no CPU fallback anywhere in this file: every failure of the CUDA runtime is reported.
#include "engine_internal.h"
#include "onee.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <numeric>
#include <thread>
#include <tuple>

using namespace qcf;

// the 21 angular classes (bra pair class >= ket pair class), one object file each
#define QCF_DECL(a, b, c, d) extern "C" const ClassLaunch qcf_class_##a##b##c##d;
#define QCF_ALL_CLASSES(X) \
    X(0, 0, 0, 0) \
    X(1, 0, 0, 0) X(1, 0, 1, 0) \
    X(1, 1, 0, 0) X(1, 1, 1, 0) X(1, 1, 1, 1) \
    X(2, 0, 0, 0) X(2, 0, 1, 0) X(2, 0, 1, 1) X(2, 0, 2, 0) \
    X(2, 1, 0, 0) X(2, 1, 1, 0) X(2, 1, 1, 1) X(2, 1, 2, 0) X(2, 1, 2, 1) \
    X(2, 2, 0, 0) X(2, 2, 1, 0) X(2, 2, 1, 1) X(2, 2, 2, 0) X(2, 2, 2, 1) X(2, 2, 2, 2)
QCF_ALL_CLASSES(QCF_DECL)

namespace {

#define CK QCF_CK

constexpr int NPAIRCLASS = 6;
inline int pair_class(int la, int lb) { return la * (la + 1) / 2 + lb; }  // ss0 ps1 pp2 ds3 dp4 dd5

const ClassLaunch* class_table(int bra_cls, int ket_cls) {
    static const ClassLaunch* tab[NPAIRCLASS][NPAIRCLASS] = {};
    static bool init = false;
    if (!init) {
#define QCF_REG(a, b, c, d) tab[pair_class(a, b)][pair_class(c, d)] = &qcf_class_##a##b##c##d;
        QCF_ALL_CLASSES(QCF_REG)
        init = true;
    }
    return tab[bra_cls][ket_cls];
}

// SURVEY.md 8d op-count model per primitive quartet of class (la lb | lc ld)
double model_flops_prim(int la, int lb, int lc, int ld) {
    const int Lab = la + lb, Lcd = lc + ld, L = Lab + Lcd;
    const int nab = ncart(la) * ncart(lb), ncd = ncart(lc) * ncart(ld);
    double f = (20 + 3 * L) + 30;
    for (int n = 0; n < L; ++n) f += 3.0 * nherm(L - n);
    f += 2.0 * nherm(Lab) * nherm(Lcd) * nab;
    f += 2.0 * nherm(Lcd) * nab * ncd;
    return f;
}

}  // namespace

// Persistent helper threads for the two staging copies of a host call (caller's pageable P -> pinned, pinned -> caller's
// G): one thread moves 8 MB at ~10 GB/s, i.e. 0.8 ms per matrix and call -- 15 % of an 11 ms build on 8 GPUs.
struct qcf_copy_pool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv, done_cv;
    std::function<void(int)> job;
    int generation = 0, pending = 0;
    bool stop = false;
    explicit qcf_copy_pool(int n) {
        for (int w = 0; w < n; ++w)
            workers.emplace_back([this, w] {
                int seen = 0;
                while (true) {
                    std::function<void(int)> j;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv.wait(lk, [&] { return stop || generation != seen; });
                        if (stop) return;
                        seen = generation;
                        j = job;
                    }
                    j(w + 1);
                    {
                        std::lock_guard<std::mutex> lk(m);
                        if (--pending == 0) done_cv.notify_all();
                    }
                }
            });
    }
    ~qcf_copy_pool() {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv.notify_all();
        for (auto& t : workers) t.join();
    }
    int threads() const { return (int)workers.size() + 1; }
    // part(i) for i = 0 .. threads()-1, part 0 on the calling thread; returns when all are done
    void run(const std::function<void(int)>& part) {
        if (workers.empty()) { part(0); return; }
        {
            std::lock_guard<std::mutex> lk(m);
            job = part; pending = (int)workers.size(); ++generation;
        }
        cv.notify_all();
        part(0);
        std::unique_lock<std::mutex> lk(m);
        done_cv.wait(lk, [&] { return pending == 0; });
    }
    // dst[0..bytes) = src[0..bytes), split over the calling thread and the workers
    void copy(void* dst, const void* src, size_t bytes) {
        const int parts = threads();
        if (parts == 1 || bytes < (1u << 20)) { std::memcpy(dst, src, bytes); return; }
        const size_t chunk = ((bytes + parts - 1) / parts + 63) & ~size_t(63);
        run([=](int i) {
            const size_t lo = std::min(bytes, chunk * i), hi = std::min(bytes, chunk * (i + 1));
            if (hi > lo) std::memcpy((char*)dst + lo, (const char*)src + lo, hi - lo);
        });
    }
};

namespace {

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ---- Boys table (host, long double) ------------------------------------------------------------------
long double boys_series(int m, long double T) {
    long double term = 1.0L / (2 * m + 1), sum = term;
    for (int k = 1; k < 600; ++k) {
        term *= 2.0L * T / (2 * m + 2 * k + 1);
        sum += term;
        if (term < 1e-24L * sum) break;
    }
    return expl(-T) * sum;
}

std::vector<double> make_boys_table() {
    const int MR = BOYS_LTOT + BOYS_ORDER + 1;
    std::vector<double> tab((size_t)(BOYS_LTOT + 1) * BOYS_NGRID * BOYS_ROW, 0.0);
    std::vector<long double> F(MR);
    for (int g = 0; g < BOYS_NGRID; ++g) {
        const long double T0 = (long double)g / BOYS_PER_UNIT;
        F[MR - 1] = boys_series(MR - 1, T0);
        const long double e = expl(-T0);
        for (int m = MR - 1; m > 0; --m) F[m - 1] = (2.0L * T0 * F[m] + e) / (2 * m - 1);
        for (int L = 0; L <= BOYS_LTOT; ++L) {
            long double fact = 1.0L;
            for (int k = 0; k <= BOYS_ORDER; ++k) {
                if (k > 0) fact *= k;
                tab[((size_t)L * BOYS_NGRID + g) * BOYS_ROW + k] = (double)(F[L + k] / fact);
            }
            tab[((size_t)L * BOYS_NGRID + g) * BOYS_ROW + BOYS_ORDER + 1] = (double)e;
        }
    }
    return tab;
}

// ---- small kernels --------------------------------------------------------------------------------
// Pk0 = fs fs Pa (and Pk1, Pj = Pk0 + Pk1 for two densities).  With `prev` buffers the build is incremental:
// the scaled matrices hold P - P_prev and P_prev is replaced by P (difference-density build).
__global__ void scale_density_kernel(int N, const double* __restrict__ fs, const double* __restrict__ Pa,
                                     const double* __restrict__ Pb, double* __restrict__ Pk0, double* __restrict__ Pk1,
                                     double* __restrict__ Pj, double* __restrict__ prev_a, double* __restrict__ prev_b) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)N * N) return;
    const int i = idx / N, j = idx % N;
    const double s = fs[i] * fs[j];
    double pa = Pa[idx];
    if (prev_a) { const double o = prev_a[idx]; prev_a[idx] = pa; pa -= o; }
    const double a = pa * s;
    Pk0[idx] = a;
    if (Pb) {
        double pb = Pb[idx];
        if (prev_b) { const double o = prev_b[idx]; prev_b[idx] = pb; pb -= o; }
        const double b = pb * s;
        Pk1[idx] = b;
        Pj[idx] = a + b;
    }
}

// one thread per shell block: max |P| over the block and over all densities
__global__ void dens_block_max_kernel(int N, int nshell, const int* __restrict__ shoff, const double* __restrict__ P0,
                                      const double* __restrict__ P1, const double* __restrict__ P2, float* __restrict__ Dsh,
                                      BuildScalars* __restrict__ sc) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    if (idx < nshell * nshell) {
        const int sa = idx / nshell, sb = idx % nshell;
        double mm = 0.0;
        for (int i = shoff[sa]; i < shoff[sa + 1]; ++i)
            for (int j = shoff[sb]; j < shoff[sb + 1]; ++j) {
                const size_t k = (size_t)i * N + j;
                mm = fmax(mm, fabs(P0[k]));
                if (P1) mm = fmax(mm, fabs(P1[k]));
                if (P2) mm = fmax(mm, fabs(P2[k]));
            }
        // round up so that the float bound never undercuts the double value
        m = __double2float_ru(mm);
        Dsh[idx] = m;
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(&sc->dmax_bits, __float_as_uint(m));
}

// per-pair density maxima of all groups (the ket scans read them coalesced instead of gathering Dsh[sc][sd])
__global__ void pair_dmax_kernel(int npair, int nshell, const int* __restrict__ sa, const int* __restrict__ sb,
                                 const float* __restrict__ Dsh, float* __restrict__ Dp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npair) Dp[i] = Dsh[(size_t)sa[i] * nshell + sb[i]];
}

// Deterministic mode: S = sum |P| in a fixed order (one block, strided partial sums, tree), then the power-of-two
// fixed-point scale 2^(61 - e) with 2^e >= 4 qmax^2 S, an upper bound of every accumulator (|I| <= Q_ab Q_cd).
__global__ void fixed_point_scale_kernel(size_t nn, const double* __restrict__ P0, const double* __restrict__ P1, double qmax,
                                         BuildScalars* __restrict__ sc) {
    __shared__ double part[1024];
    double s = 0.0;
    for (size_t i = threadIdx.x; i < nn; i += 1024) {
        s += fabs(P0[i]);
        if (P1) s += fabs(P1[i]);
    }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double S = part[0];
        const double B = 4.0 * qmax * qmax * S;
        int e = 0;
        if (B > 0.0) frexp(B, &e);
        sc->abs_sum = S;
        sc->fx_scale = ldexp(1.0, 61 - e);
    }
}

// accumulators of up to QCF_MAXDEV devices (peer pointers over NVLink when ndev > 1)
struct PeerAcc {
    const double* AJ[QCF_MAXDEV];
    const double* AK0[QCF_MAXDEV];
    const double* AK1[QCF_MAXDEV];
};

// A[idx] + A[tr] summed over the devices in FIXED order; integer arithmetic in the deterministic mode
__device__ __forceinline__ double sym_sum(const double* const* A, int ndev, size_t idx, size_t tr, double fx, double inv_fx) {
    if (fx != 0.0) {
        long long t = 0;
        for (int d = 0; d < ndev; ++d) {
            const long long* p = reinterpret_cast<const long long*>(A[d]);
            t += p[idx] + p[tr];
        }
        return (double)t * inv_fx;
    }
    double t = 0.0;
    for (int d = 0; d < ndev; ++d) t += A[d][idx] + A[d][tr];
    return t;
}

// Symmetrise, apply the component scales, combine J and K, and -- when the context drives several GPUs -- sum the
// partial accumulators of all devices through peer loads (this device produces rows [row0, row1) of the result and
// stores them into dev[0]'s G: a fused finalize + reduce-scatter + gather over NVLink, no staging copies).
// mode 0: G0 = fs fs (2(AJ+AJ^T) - 1/2 (AK0+AK0^T))                       (RHF)
// mode 1: G0/G1 = fs fs (2(AJ+AJ^T) - (AKs + AKs^T))                      (UHF)
// mode 2: G0 = fs fs 2(AJ+AJ^T),  G1 = fs fs (AK0+AK0^T)                  (J and K)
// acc != null: G += result (incremental builds)
__global__ void finalize_kernel(int N, int mode, int ndev, PeerAcc acc, const double* __restrict__ fs,
                                const BuildScalars* __restrict__ sc, int row0, int row1, double* __restrict__ G0,
                                double* __restrict__ G1, int accumulate) {
    const size_t idx = (size_t)row0 * N + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)row1 * N) return;
    const int i = idx / N, j = idx % N;
    const size_t tr = (size_t)j * N + i;
    const double fx = sc->fx_scale, inv_fx = fx != 0.0 ? 1.0 / fx : 0.0;
    const double s = fs[i] * fs[j];
    const double J = 2.0 * sym_sum(acc.AJ, ndev, idx, tr, fx, inv_fx);
    const double K0 = sym_sum(acc.AK0, ndev, idx, tr, fx, inv_fx);
    double g0, g1 = 0.0;
    if (mode == 0) {
        g0 = s * (J - 0.5 * K0);
    } else if (mode == 1) {
        const double K1 = sym_sum(acc.AK1, ndev, idx, tr, fx, inv_fx);
        g0 = s * (J - K0);
        g1 = s * (J - K1);
    } else {
        g0 = s * J;
        g1 = s * K0;
    }
    if (accumulate) {
        G0[idx] += g0;
        if (mode != 0) G1[idx] += g1;
    } else {
        G0[idx] = g0;
        if (mode != 0) G1[idx] = g1;
    }
}

template <int L>
__global__ void boys_test_kernel(int n, const double* __restrict__ T, const double* __restrict__ table, double* __restrict__ F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double f[L + 1];
    boys<L>(T[i], table, f);
#pragma unroll
    for (int m = 0; m <= L; ++m) F[(size_t)i * (L + 1) + m] = f[m];
}

// FP64 FMA peak: 8 independent dependent chains per thread
__global__ void fp64_peak_kernel(double* out, int iters, double x) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double y = x * 0.5;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
        a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}

// ---- pair construction -------------------------------------------------------------------------------
struct PairArrays {
    std::vector<int> fa, fb, sa, sb, np;
    std::vector<double> Q, Qb, prim, AB;
};

// power-of-two ceiling of a Schwarz factor: Q <= 2^ceil(log2 Q)
inline int q_bucket(double Q) { return Q > 0 ? (int)std::ceil(std::log2(Q)) : -100000; }
inline double q_bucket_ceiling(double Q) { return Q > 0 ? std::ldexp(1.0, q_bucket(Q)) : 0.0; }

void fill_pair_arrays(const qcf_ctx* c, const HostGroup& g, PairArrays& out) {
    const size_t np = g.pairs.size();
    out.fa.resize(np); out.fb.resize(np); out.sa.resize(np); out.sb.resize(np); out.Q.resize(np); out.Qb.resize(np); out.np.resize(np);
    out.prim.assign((size_t)g.K * PF_COUNT * np, 0.0);
    out.AB.assign(3 * np, 0.0);
    const double cpi = std::sqrt(2.0) * std::pow(PI_D, 1.25);
    for (size_t i = 0; i < np; ++i) {
        const int sa = g.pairs[i].sa, sb = g.pairs[i].sb;
        out.fa[i] = c->sh_off[sa]; out.fb[i] = c->sh_off[sb]; out.sa[i] = sa; out.sb[i] = sb; out.Q[i] = g.pairs[i].Q; out.Qb[i] = q_bucket_ceiling(g.pairs[i].Q);
        const double* A = &c->xyz[3 * c->sh_atom[sa]];
        const double* B = &c->xyz[3 * c->sh_atom[sb]];
        double AB2 = 0;
        for (int k = 0; k < 3; ++k) { out.AB[k * np + i] = A[k] - B[k]; AB2 += (A[k] - B[k]) * (A[k] - B[k]); }
        out.np[i] = g.pairs[i].keff < 0 ? g.K : g.pairs[i].keff;
        const int npb = c->sh_np[sb];
        for (int kk = 0; kk < g.K; ++kk) {
            const int ksel = g.pairs[i].order.empty() ? kk : g.pairs[i].order[kk];
            const int ia = ksel / npb, ib = ksel % npb;
            const double a = c->exps[c->sh_po[sa] + ia], b = c->exps[c->sh_po[sb] + ib];
            const double ca = c->coefs[c->sh_po[sa] + ia], cb = c->coefs[c->sh_po[sb] + ib];
            const double p = a + b, mu = a * b / p;
            double* f = &out.prim[(size_t)kk * PF_COUNT * np + i];
            f[PF_P * np] = p;
            for (int k = 0; k < 3; ++k) {
                const double Pk = (a * A[k] + b * B[k]) / p;
                f[(PF_PX + k) * np] = Pk;
                f[(PF_PAX + k) * np] = Pk - A[k];
            }
            f[PF_C * np] = cpi * ca * cb * std::exp(-mu * AB2) / p;
        }
    }
}

template <class T>
cudaError_t upload(T** dptr, const std::vector<T>& h) {
    cudaError_t e = cudaMalloc((void**)dptr, std::max<size_t>(h.size(), 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (h.empty()) return cudaSuccess;
    return cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

void free_group(GroupDev& g) {
    cudaFree(g.fa); cudaFree(g.fb); cudaFree(g.sa); cudaFree(g.sb); cudaFree(g.np); cudaFree(g.Q); cudaFree(g.Qb); cudaFree(g.prim); cudaFree(g.AB);
    cudaFree(g.bra_list);
    g = GroupDev{};
}

// upload one group's pair data to the current device
int upload_group(qcf_ctx* ctx, const HostGroup& g, GroupDev& d) {
    PairArrays pa;
    fill_pair_arrays(ctx, g, pa);
    free_group(d);
    CK(upload(&d.fa, pa.fa)); CK(upload(&d.fb, pa.fb)); CK(upload(&d.sa, pa.sa)); CK(upload(&d.sb, pa.sb)); CK(upload(&d.np, pa.np));
    CK(upload(&d.Q, pa.Q)); CK(upload(&d.Qb, pa.Qb)); CK(upload(&d.prim, pa.prim)); CK(upload(&d.AB, pa.AB));
    d.pg.npair = (int)g.pairs.size(); d.pg.K = g.K; d.pg.la = g.la; d.pg.lb = g.lb;
    d.pg.fa = d.fa; d.pg.fb = d.fb; d.pg.sa = d.sa; d.pg.sb = d.sb; d.pg.nprim = d.np; d.pg.Q = d.Q; d.pg.Qb = d.Qb; d.pg.prim = d.prim; d.pg.AB = d.AB;
    d.pg.Dp = nullptr;
    d.nbra = d.pg.npair;
    return QCF_OK;
}

// Schwarz factors of one host group on the current device (dev[0]); the group may be a temporary K = 1 list
int schwarz_of(qcf_ctx* ctx, const HostGroup& g, std::vector<double>& Q) {
    GroupDev d;
    int rc = upload_group(ctx, g, d);
    if (rc) return rc;
    const int np = (int)g.pairs.size();
    double* dQ = nullptr;
    CK(cudaMalloc(&dQ, sizeof(double) * std::max(np, 1)));
    class_table(g.cls, g.cls)->schwarz((np + 63) / 64, 64, 0, d.pg, ctx->dev[0].boys, dQ);
    CK(cudaGetLastError());
    Q.resize(np);
    CK(cudaMemcpy(Q.data(), dQ, sizeof(double) * np, cudaMemcpyDeviceToHost));
    cudaFree(dQ);
    free_group(d);
    return QCF_OK;
}

// host-side pair lists: significant pairs, Schwarz factors, primitive screening, sorting (device work on dev[0])
int build_pairs(qcf_ctx* ctx) {
    const int ns = ctx->nshell;
    std::map<std::tuple<int, int>, int> gid;   // (cls, K) -> group
    for (int s1 = 0; s1 < ns; ++s1)
        for (int s2 = 0; s2 <= s1; ++s2) {
            int sa = s1, sb = s2;
            if (ctx->sh_l[sb] > ctx->sh_l[sa]) std::swap(sa, sb);
            const int la = ctx->sh_l[sa], lb = ctx->sh_l[sb];
            if (ctx->screening) {   // distance / overlap prescreen
                const double* A = &ctx->xyz[3 * ctx->sh_atom[sa]];
                const double* B = &ctx->xyz[3 * ctx->sh_atom[sb]];
                double R2 = 0;
                for (int k = 0; k < 3; ++k) R2 += (A[k] - B[k]) * (A[k] - B[k]);
                double est = 0;
                for (int ia = 0; ia < ctx->sh_np[sa]; ++ia)
                    for (int ib = 0; ib < ctx->sh_np[sb]; ++ib) {
                        const double a = ctx->exps[ctx->sh_po[sa] + ia], b = ctx->exps[ctx->sh_po[sb] + ib];
                        const double p = a + b;
                        est = std::max(est, std::fabs(ctx->coefs[ctx->sh_po[sa] + ia] * ctx->coefs[ctx->sh_po[sb] + ib]) *
                                                std::pow(PI_D / p, 1.5) * std::exp(-a * b / p * R2));
                    }
                est *= std::pow(1.0 + std::sqrt(R2), la + lb);
                if (est < 1e-18) continue;
            }
            const int K = ctx->sh_np[sa] * ctx->sh_np[sb];
            const int cls = pair_class(la, lb);
            auto key = std::make_tuple(cls, K);
            auto it = gid.find(key);
            if (it == gid.end()) {
                HostGroup g; g.la = la; g.lb = lb; g.K = K; g.cls = cls;
                ctx->groups.push_back(std::move(g));
                it = gid.emplace(key, (int)ctx->groups.size() - 1).first;
            }
            ctx->groups[it->second].pairs.push_back({sa, sb, 0.0});
        }
    std::sort(ctx->groups.begin(), ctx->groups.end(), [](const HostGroup& x, const HostGroup& y) {
        return x.cls != y.cls ? x.cls < y.cls : x.K < y.K;
    });
    // Schwarz factors on the device, group by group
    ctx->qmax = 0;
    for (auto& g : ctx->groups) {
        std::vector<double> Q;
        int rc = schwarz_of(ctx, g, Q);
        if (rc) return rc;
        for (size_t i = 0; i < g.pairs.size(); ++i) { g.pairs[i].Q = Q[i]; ctx->qmax = std::max(ctx->qmax, Q[i]); }
    }
    // primitive screening: Schwarz factor of every primitive pair on its own (the same class kernel on a
    // K = 1 list); inside each shell pair the primitives are sorted by it and those that cannot contribute
    // Q_k * Q_max >= tau to any integral are dropped (the kernels loop over nprim[i] <= K primitives).  Measured at
    // N = 1007 (profiles/r2_ab_call8_prim_cut.log, r2_ab_call11_cuts.log): factor 1e-4 / 1e-1 / 1 / 10 keeps 66 / 59 / 56 / 53 %
    // of the primitive pairs, 70.8 / 69.2 / 68.2 / 67.1 ms, max |dG| vs the unscreened oracle 3.7e-11 / 3.7e-11 / 3.7e-11 / 4.6e-11
    ctx->prim_total = ctx->prim_kept = 0;
    for (auto& g : ctx->groups) {
        const long long np = (long long)g.pairs.size();
        ctx->prim_total += np * g.K;
        if (!ctx->screening || g.K == 1) { ctx->prim_kept += np * g.K; continue; }
        HostGroup g1; g1.la = g.la; g1.lb = g.lb; g1.K = 1; g1.cls = g.cls;
        g1.pairs.reserve((size_t)np * g.K);
        for (const auto& pr : g.pairs)
            for (int k = 0; k < g.K; ++k) { HostPair h{pr.sa, pr.sb, 0.0}; h.order.assign(1, (uint16_t)k); g1.pairs.push_back(std::move(h)); }
        // the K = 1 list holds primitive `order[0]` of each pair in slot 0
        std::vector<double> Qk;
        int rc = schwarz_of(ctx, g1, Qk);
        if (rc) return rc;
        const double pcut = ctx->tau * ctx->prim_cut_factor / std::max(ctx->qmax, 1e-300);
        for (long long i = 0; i < np; ++i) {
            auto& pr = g.pairs[i];
            pr.order.resize(g.K);
            for (int k = 0; k < g.K; ++k) pr.order[k] = (uint16_t)k;
            const double* q = &Qk[(size_t)i * g.K];
            std::stable_sort(pr.order.begin(), pr.order.end(), [&](uint16_t x, uint16_t y) { return q[x] > q[y]; });
            int keep = 0;
            while (keep < g.K && q[pr.order[keep]] >= pcut) ++keep;
            pr.keff = std::max(keep, 1);
            ctx->prim_kept += pr.keff;
        }
    }
    // drop negligible pairs, sort by Q bucket
    size_t npairs = 0;
    for (auto& g : ctx->groups) {
        if (ctx->screening) {
            const double cut = ctx->tau * ctx->pair_cut_factor / std::max(ctx->qmax, 1e-300);
            g.pairs.erase(std::remove_if(g.pairs.begin(), g.pairs.end(), [&](const HostPair& p) { return p.Q < cut; }), g.pairs.end());
        }
        std::stable_sort(g.pairs.begin(), g.pairs.end(), [](const HostPair& x, const HostPair& y) {
            const int bx = q_bucket(x.Q), by = q_bucket(y.Q);
            if (bx != by) return bx > by;
            return x.sa != y.sa ? x.sa < y.sa : x.sb < y.sb;
        });
    }
    ctx->groups.erase(std::remove_if(ctx->groups.begin(), ctx->groups.end(), [](const HostGroup& g) { return g.pairs.empty(); }),
                      ctx->groups.end());
    for (size_t gi = 0; gi < ctx->groups.size(); ++gi) {
        auto& g = ctx->groups[gi];
        g.pair_off = npairs;
        npairs += g.pairs.size();
        for (size_t i = 0; i < g.pairs.size(); ++i) ctx->pair_index[{g.pairs[i].sa, g.pairs[i].sb}] = {(int)gi, (int)i};
    }
    ctx->npairs = npairs;
    ctx->stats.n_pairs = (int)npairs;
    ctx->stats.n_groups = (int)ctx->groups.size();
    ctx->stats.prim_pairs = ctx->prim_total;
    ctx->stats.prim_pairs_kept = ctx->prim_kept;
    return QCF_OK;
}

// Launch-shape defaults by the number of shell pairs one rank works on (measured, profiles/r2_ab_call10_rank_share_knobs.log,
// r2_ab_call18_small_problems.log): a big share fills the GPU launch by launch -- 8 streams, ~296 CTAs per launch and 64 kets per
// thread are best; the smaller the share, the more its 1/world-size (or small-molecule) grids depend on concurrency and on
// fewer, longer CTAs: 16 streams / 148 CTAs, then 32 streams / 74 CTAs / 32 kets per thread.  Environment variables override.
void tune_launch_shape(qcf_ctx* ctx) {
    const size_t share = ctx->npairs / (size_t)std::max(ctx->world, 1);
    int streams = 8, ctas = 296, kpt = ctx->world > 1 ? 32 : 64;
    if (share < 16000) { streams = 32; ctas = 74; kpt = 32; }
    else if (share < 40000) { streams = 16; ctas = 148; }
    // Bra split: per launch from four ranks on -- a launch is split over all ranks only if its bra list has at least 1184
    // pairs (eight per SM), a shorter one is shared by proportionally fewer ranks (148 pairs per rank at eight ranks, 296 at
    // four).  Measured rank by rank on one GPU for the N = 1007 build (profiles/r2_ab_call24..., r2_ab_call25...): slowest of
    // eight ranks 9.83 -> 9.42 ms, slowest of four 19.1 -> 18.0 ms.  Per group below four ranks.
    if (!getenv("QCF_SPLIT_MIN_BRAS")) ctx->split_min_bras = ctx->world >= 4 ? std::max(1, 1184 / ctx->world) : 0;
    if (!getenv("QCF_STREAMS")) ctx->nstreams = streams;
    if (!getenv("QCF_TARGET_CTAS")) ctx->target_ctas = ctas;
    if (!getenv("QCF_KETS_PER_THREAD")) ctx->kets_per_thread = kpt;
}

// ---- launch plan and the cost-balanced bra split ------------------------------------------------------
double per_quartet_cost(const HostGroup& bra, const HostGroup& ket) {
    return (double)bra.K * ket.K * model_flops_prim(bra.la, bra.lb, ket.la, ket.lb) + 400.0;
}

// modelled cost of bra pair i of group `bra` against ket group `ket` (length of its Schwarz prefix x cost per quartet)
double bra_item_cost(const qcf_ctx* ctx, const HostGroup& bra, const HostGroup& ket, int i, bool same_group) {
    int n = (int)ket.pairs.size();
    if (ctx->screening) {
        const double need = ctx->tau / std::max(bra.pairs[i].Q, 1e-300);
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (q_bucket_ceiling(ket.pairs[mid].Q) >= need) lo = mid + 1; else hi = mid; }
        n = lo;
    }
    if (same_group) n = std::min(n, i + 1);
    return (double)n * per_quartet_cost(bra, ket);
}

// Per-launch split (see qcf_ctx::launch_split).  Launches in descending modelled cost; one whose bra list gives fewer than
// split_min_bras pairs per rank goes to the s = nbra / split_min_bras currently least loaded ranks (its bra pairs heaviest
// first onto the least loaded of those); the launches long enough for every rank follow as one heaviest-first pass over
// all their (launch, bra) items, which evens out what the short ones left.
void split_per_launch(qcf_ctx* ctx) {
    const int W = ctx->world, nl = (int)ctx->plan.size();
    ctx->launch_split.assign(nl, std::vector<std::vector<int>>(W));
    std::vector<double> load(W, 0.0);
    std::vector<int> order(nl);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return ctx->plan[x].cost > ctx->plan[y].cost; });
    struct Item { int l, i; double cost; };
    std::vector<Item> wide;
    std::vector<Item> items;
    for (int l : order) {
        const PlannedLaunch& pl = ctx->plan[l];
        const HostGroup& bra = ctx->groups[pl.gi];
        const HostGroup& ket = ctx->groups[pl.gj];
        const int nbra = (int)bra.pairs.size();
        items.clear();
        for (int i = 0; i < nbra; ++i) items.push_back({l, i, bra_item_cost(ctx, bra, ket, i, pl.gi == pl.gj) + 100.0});
        const int s = std::max(1, nbra / ctx->split_min_bras);
        if (s >= W) { wide.insert(wide.end(), items.begin(), items.end()); continue; }
        std::vector<int> ranks(W);
        std::iota(ranks.begin(), ranks.end(), 0);
        std::stable_sort(ranks.begin(), ranks.end(), [&](int x, int y) { return load[x] < load[y]; });
        ranks.resize(s);
        std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.cost > y.cost; });
        for (const Item& it : items) {
            int r = ranks[0];
            for (int q : ranks) if (load[q] < load[r]) r = q;
            load[r] += it.cost;
            ctx->launch_split[l][r].push_back(it.i);
        }
    }
    std::stable_sort(wide.begin(), wide.end(), [](const Item& x, const Item& y) { return x.cost > y.cost; });
    for (const Item& it : wide) {
        const int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[r] += it.cost;
        ctx->launch_split[it.l][r].push_back(it.i);
    }
    for (auto& l : ctx->launch_split)
        for (auto& v : l) std::sort(v.begin(), v.end());
    ctx->rank_cost = load;
}

void make_plan(qcf_ctx* ctx) {
    const int ng = (int)ctx->groups.size();
    ctx->plan.clear();
    ctx->launch_split.clear();
    // Order: the launches whose threads run longest go first (they overlap with everything else), round-robin over the
    // streams.  ps: lanes per shell quartet for the highly contracted launches.
    for (int gi = ng - 1; gi >= 0; --gi)
        for (int gj = gi; gj >= 0; --gj) {
            const HostGroup& bra = ctx->groups[gi];
            const HostGroup& ket = ctx->groups[gj];
            const int nprimq = bra.K * ket.K;
            int ps = nprimq >= 9 * ctx->ps_min_prim ? 8 : (nprimq >= ctx->ps_min_prim ? 4 : 1);
            ps = std::min(ps, class_table(bra.cls, ket.cls)->max_ps);
            const double pq = per_quartet_cost(bra, ket);
            const double nq = (double)bra.pairs.size() * ket.pairs.size() * (gi == gj ? 0.5 : 1.0);
            ctx->plan.push_back({gi, gj, 0, ps, pq / ps, nq * pq});
        }
    if (ctx->launch_order == 1)         // biggest launches first
        std::stable_sort(ctx->plan.begin(), ctx->plan.end(), [](const PlannedLaunch& x, const PlannedLaunch& y) { return x.cost > y.cost; });
    else                                // default: longest-running threads first, then biggest
        std::stable_sort(ctx->plan.begin(), ctx->plan.end(), [](const PlannedLaunch& x, const PlannedLaunch& y) {
            return x.serial != y.serial ? x.serial > y.serial : x.cost > y.cost;
        });
    // Cost-balanced static split of every group's bra list over the ranks: modelled cost of bra pair i =
    // sum over the ket groups it is paired with of (length of its Schwarz prefix at a nominal density maximum of 1) x
    // (primitive quartets x op count of the class); heaviest first onto the least loaded rank, one global load
    // vector for all groups so that a surplus in one class is made up in another.
    const int W = ctx->world;
    ctx->bra_split.assign(ng, std::vector<std::vector<int>>(W));
    ctx->rank_cost.assign(W, 0.0);
    if (W == 1) {
        for (int gi = 0; gi < ng; ++gi) {
            auto& v = ctx->bra_split[gi][0];
            v.resize(ctx->groups[gi].pairs.size());
            std::iota(v.begin(), v.end(), 0);
        }
        return;
    }
    struct Item { int gi, i; double cost; };
    std::vector<Item> items;
    items.reserve(ctx->npairs);
    for (int gi = 0; gi < ng; ++gi) {
        const HostGroup& bra = ctx->groups[gi];
        for (int i = 0; i < (int)bra.pairs.size(); ++i) {
            double c = 0;
            for (int gj = 0; gj <= gi; ++gj) {
                const HostGroup& ket = ctx->groups[gj];
                int n = (int)ket.pairs.size();
                if (ctx->screening) {
                    const double need = ctx->tau / std::max(bra.pairs[i].Q, 1e-300);
                    int lo = 0, hi = n;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (q_bucket_ceiling(ket.pairs[mid].Q) >= need) lo = mid + 1; else hi = mid; }
                    n = lo;
                }
                if (gj == gi) n = std::min(n, i + 1);
                c += (double)n * per_quartet_cost(bra, ket);
            }
            items.push_back({gi, i, c + 2000.0});   // + the CTA prologues of one bra pair
        }
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.cost > y.cost; });
    for (const Item& it : items) {
        const int r = (int)(std::min_element(ctx->rank_cost.begin(), ctx->rank_cost.end()) - ctx->rank_cost.begin());
        ctx->rank_cost[r] += it.cost;
        ctx->bra_split[it.gi][r].push_back(it.i);
    }
    for (auto& g : ctx->bra_split)
        for (auto& v : g) std::sort(v.begin(), v.end());
    if (ctx->split_min_bras > 0) {
        // keep the per-launch split only if its items are fine enough to balance (small molecules: a few hundred whole
        // launches cannot be spread evenly over eight ranks; the per-group split always can)
        const std::vector<double> group_cost = ctx->rank_cost;
        split_per_launch(ctx);
        double cmax = 0, csum = 0;
        for (double c : ctx->rank_cost) { cmax = std::max(cmax, c); csum += c; }
        if (!(csum > 0) || cmax > 1.01 * csum / W) { ctx->launch_split.clear(); ctx->rank_cost = group_cost; }
    }
}

// ---- per-device state ---------------------------------------------------------------------------------
int setup_device(qcf_ctx* ctx, qcf_device& dv, const std::vector<double>& boys_tab) {
    const size_t nn = (size_t)ctx->N * ctx->N;
    CK(cudaSetDevice(dv.device));
    if (!dv.boys) CK(upload(&dv.boys, boys_tab));
    CK(upload(&dv.fscale, ctx->fscale));
    CK(upload(&dv.shoff, ctx->sh_off));
    CK(cudaStreamCreateWithFlags(&dv.main, cudaStreamNonBlocking));
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));     // numerically lower = higher priority
    for (int s = 0; s < QCF_MAXSTREAM; ++s) {
        // QCF_PRIO: odd streams carry one kernel family at the other priority (see enqueue_device_work)
        const int prio = (ctx->stream_prio != 0 && (s & 1)) ? prio_lo : prio_hi;
        CK(cudaStreamCreateWithPriority(&dv.streams[s], cudaStreamNonBlocking, ctx->stream_prio != 0 ? prio : 0));
        CK(cudaEventCreateWithFlags(&dv.ev_join[s], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&dv.ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&dv.ev_acc, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&dv.ev_red, cudaEventDisableTiming));
    CK(cudaEventCreate(&dv.ev_k0)); CK(cudaEventCreate(&dv.ev_k1));
    for (int k = 0; k < 2; ++k) {
        CK(cudaMalloc(&dv.Pin[k], nn * sizeof(double)));
        CK(cudaMalloc(&dv.Pk[k], nn * sizeof(double)));
        CK(cudaMalloc(&dv.AK[k], nn * sizeof(double)));
        CK(cudaMalloc(&dv.G[k], nn * sizeof(double)));
    }
    CK(cudaMalloc(&dv.Pj, nn * sizeof(double)));
    CK(cudaMalloc(&dv.AJ, nn * sizeof(double)));
    CK(cudaMalloc(&dv.Dsh, sizeof(float) * ctx->nshell * ctx->nshell));
    CK(cudaMalloc(&dv.sc, sizeof(BuildScalars)));
    CK(cudaMemset(dv.sc, 0, sizeof(BuildScalars)));
    CK(cudaMalloc(&dv.counters, sizeof(unsigned long long) * std::max<size_t>(ctx->plan.size(), 1)));
    // pair data of every group + concatenated shell ids / density maxima
    const int ng = (int)ctx->groups.size();
    dv.groups.assign(ng, GroupDev{});
    std::vector<int> all_sa(ctx->npairs), all_sb(ctx->npairs);
    CK(cudaMalloc(&dv.all_Dp, sizeof(float) * std::max<size_t>(ctx->npairs, 1)));
    CK(cudaMemset(dv.all_Dp, 0, sizeof(float) * std::max<size_t>(ctx->npairs, 1)));
    int max_bra_K[NPAIRCLASS] = {};
    for (int gi = 0; gi < ng; ++gi) {
        const HostGroup& g = ctx->groups[gi];
        int rc = upload_group(ctx, g, dv.groups[gi]);
        if (rc) return rc;
        dv.groups[gi].pg.Dp = dv.all_Dp + g.pair_off;
        for (size_t i = 0; i < g.pairs.size(); ++i) { all_sa[g.pair_off + i] = g.pairs[i].sa; all_sb[g.pair_off + i] = g.pairs[i].sb; }
        max_bra_K[g.cls] = std::max(max_bra_K[g.cls], g.K);
        const auto& list = ctx->bra_split[gi][dv.rank];
        dv.groups[gi].nbra = (int)list.size();
        if (ctx->world > 1) CK(upload(&dv.groups[gi].bra_list, list));
    }
    CK(upload(&dv.all_sa, all_sa));
    CK(upload(&dv.all_sb, all_sb));
    if (!ctx->launch_split.empty()) {
        dv.launch_list.assign(ctx->plan.size(), nullptr);
        dv.launch_nbra.assign(ctx->plan.size(), 0);
        for (size_t l = 0; l < ctx->plan.size(); ++l) {
            const auto& list = ctx->launch_split[l][dv.rank];
            dv.launch_nbra[l] = (int)list.size();
            if (!list.empty()) CK(upload(&dv.launch_list[l], list));
        }
    }
    // opt in to large dynamic shared memory once per device and kernel (not per launch)
    for (int b = 0; b < NPAIRCLASS; ++b)
        for (int k = 0; k <= b; ++k)
            if (max_bra_K[b] > 0) CK(class_table(b, k)->init(max_bra_K[b], ctx->nshell));
    return QCF_OK;
}

void destroy_device(qcf_device& dv) {
    cudaSetDevice(dv.device);
    cudaDeviceSynchronize();
    for (auto& g : dv.groups) free_group(g);
    for (int* p : dv.launch_list) cudaFree(p);
    if (dv.graph) cudaGraphExecDestroy(dv.graph);
    cudaFree(dv.boys); cudaFree(dv.fscale); cudaFree(dv.shoff); cudaFree(dv.all_sa); cudaFree(dv.all_sb); cudaFree(dv.all_Dp);
    for (int k = 0; k < 2; ++k) {
        cudaFree(dv.Pin[k]); cudaFree(dv.Pk[k]); cudaFree(dv.AK[k]); cudaFree(dv.G[k]); cudaFree(dv.Pprev[k]); cudaFree(dv.Gprev[k]);
    }
    cudaFree(dv.Pj); cudaFree(dv.AJ); cudaFree(dv.Dsh); cudaFree(dv.sc); cudaFree(dv.counters);
    for (int s = 0; s < QCF_MAXSTREAM; ++s) { if (dv.streams[s]) cudaStreamDestroy(dv.streams[s]); if (dv.ev_join[s]) cudaEventDestroy(dv.ev_join[s]); }
    if (dv.main) cudaStreamDestroy(dv.main);
    for (cudaEvent_t e : {dv.ev_fork, dv.ev_k0, dv.ev_k1, dv.ev_acc, dv.ev_red}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : dv.prof_ev) cudaEventDestroy(e);
}

// ---- the build ---------------------------------------------------------------------------------------
// Everything one device does up to (not including) the finalize step, enqueued on `ms` and the device's side
// streams.  pa/pb: unscaled densities on this device.  Capturable: no host synchronisation, no allocation.
int enqueue_device_work(qcf_ctx* ctx, qcf_device& dv, int mode, const double* pa, const double* pb, bool incremental, cudaStream_t ms) {
    const int N = ctx->N, ns = ctx->nshell;
    const size_t nn = (size_t)N * N;
    const int nk = (mode == 1) ? 2 : 1;
    const int tpb = 256;
    const int nblk = (int)((nn + tpb - 1) / tpb);
    CK(cudaMemsetAsync(dv.sc, 0, sizeof(BuildScalars), ms));
    scale_density_kernel<<<nblk, tpb, 0, ms>>>(N, dv.fscale, pa, nk == 2 ? pb : nullptr, dv.Pk[0], dv.Pk[1], dv.Pj,
                                               incremental ? dv.Pprev[0] : nullptr, incremental && nk == 2 ? dv.Pprev[1] : nullptr);
    const double* Pj = nk == 2 ? dv.Pj : dv.Pk[0];
    dens_block_max_kernel<<<(ns * ns + 255) / 256, 256, 0, ms>>>(N, ns, dv.shoff, dv.Pk[0], nk == 2 ? dv.Pk[1] : nullptr,
                                                                 nk == 2 ? dv.Pj : nullptr, dv.Dsh, dv.sc);
    pair_dmax_kernel<<<(int)((ctx->npairs + 255) / 256), 256, 0, ms>>>((int)ctx->npairs, ns, dv.all_sa, dv.all_sb, dv.Dsh, dv.all_Dp);
    if (ctx->deterministic)
        fixed_point_scale_kernel<<<1, 1024, 0, ms>>>(nn, dv.Pk[0], nk == 2 ? dv.Pk[1] : nullptr, ctx->qmax, dv.sc);
    CK(cudaMemsetAsync(dv.AJ, 0, nn * sizeof(double), ms));
    CK(cudaMemsetAsync(dv.AK[0], 0, nn * sizeof(double), ms));
    if (nk == 2) CK(cudaMemsetAsync(dv.AK[1], 0, nn * sizeof(double), ms));
    CK(cudaMemsetAsync(dv.counters, 0, sizeof(unsigned long long) * std::max<size_t>(ctx->plan.size(), 1), ms));
    BuildArgs a{};
    a.N = N; a.nshell = ns; a.nk = nk;
    a.Pj = Pj; a.Pk0 = dv.Pk[0]; a.Pk1 = dv.Pk[1];
    a.AJ = dv.AJ; a.AK0 = dv.AK[0]; a.AK1 = dv.AK[1];
    a.Dsh = dv.Dsh;
    // difference-density builds: the errors of up to `full_every` consecutive screened builds add up in G, so each one
    // is screened an order of magnitude tighter
    a.tau = ctx->screening ? ctx->tau * (incremental ? 0.125 : 1.0) : 0.0;
    a.red_eps = ctx->red_eps_factor * a.tau;
    a.sc = dv.sc;
    a.boys = dv.boys;

    const int nstr = ctx->profile ? 1 : ctx->nstreams;
    CK(cudaEventRecord(dv.ev_fork, ms));
    for (int s = 0; s < nstr; ++s) CK(cudaStreamWaitEvent(dv.streams[s], dv.ev_fork, 0));
    // Kets per thread: long lists amortise the per-CTA prologue / J_ab reduction over many kets, short lists are cut
    // finer so that the grid still fills the 148 SMs (d-bra classes, a rank's share of a multi-GPU run), and the serial
    // work of one thread is capped so that highly contracted classes do not become a latency tail.
    int nl = 0, launched = 0;
    for (const PlannedLaunch& pl : ctx->plan) {
        const HostGroup& bra = ctx->groups[pl.gi];
        const HostGroup& ket = ctx->groups[pl.gj];
        const GroupDev& db = dv.groups[pl.gi];
        const GroupDev& dk = dv.groups[pl.gj];
        const int idx = nl++;
        const bool per_launch = !dv.launch_nbra.empty();
        const int nbra = per_launch ? dv.launch_nbra[idx] : db.nbra;
        if (nbra <= 0) continue;
        const int nket_max = dk.pg.npair;
        const bool slab = bra.la == 2 && bra.lb >= 1;
        const int cta_threads = slab ? 128 : ctx->block;
        const long long want_chunks = (ctx->target_ctas + nbra - 1) / nbra;
        int kpt = (int)(nket_max / (want_chunks * cta_threads));
        kpt = std::min(kpt, (int)(ctx->serial_cap / pl.serial));
        kpt = std::max(1, std::min(kpt, ctx->kets_per_thread));
        BuildArgs al = a;
        al.counter = dv.counters + idx;
        al.bra_list = per_launch ? dv.launch_list[idx] : db.bra_list;
        int si = ctx->profile ? 0 : (launched % nstr);
        if (ctx->stream_prio != 0 && !ctx->profile && nstr >= 2) {
            // even streams: high priority, odd streams: low priority.  stream_prio 1: block kernels high, slab low; 2: reverse
            const bool high = (ctx->stream_prio == 1) ? !slab : slab;
            si = (si & ~1) | (high ? 0 : 1);
        }
        cudaStream_t st = dv.streams[si];
        if (ctx->profile) {
            while ((int)dv.prof_ev.size() < 2 * (idx + 1)) { cudaEvent_t e; CK(cudaEventCreate(&e)); dv.prof_ev.push_back(e); }
            CK(cudaEventRecord(dv.prof_ev[2 * idx], st));
        }
        class_table(bra.cls, ket.cls)->jk(nk, pl.ps, nbra, nket_max, ctx->block, kpt, st, db.pg, dk.pg, al, pl.gi == pl.gj ? 1 : 0);
        if (ctx->profile) CK(cudaEventRecord(dv.prof_ev[2 * idx + 1], st));
        ++launched;
    }
    CK(cudaGetLastError());
    for (int s = 0; s < nstr; ++s) {
        CK(cudaEventRecord(dv.ev_join[s], dv.streams[s]));
        CK(cudaStreamWaitEvent(ms, dv.ev_join[s], 0));
    }
    dv.launches = launched + 3 + (ctx->deterministic ? 1 : 0);
    return QCF_OK;
}

// the device's share of a build on stream `ms`: replay the captured graph (re-captured when mode or input pointers change)
int launch_device_work(qcf_ctx* ctx, qcf_device& dv, int mode, const double* pa, const double* pb, bool incremental, cudaStream_t ms, int* graph_launches) {
    CK(cudaSetDevice(dv.device));
    const int gmode = mode + (incremental ? 8 : 0);
    if (!ctx->use_graph || ctx->profile) return enqueue_device_work(ctx, dv, mode, pa, pb, incremental, ms);
    if (!dv.graph || dv.graph_mode != gmode || dv.graph_pa != pa || dv.graph_pb != pb) {
        if (dv.graph) { cudaGraphExecDestroy(dv.graph); dv.graph = nullptr; }
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(dv.main, cudaStreamCaptureModeRelaxed));
        int rc = enqueue_device_work(ctx, dv, mode, pa, pb, incremental, dv.main);
        cudaError_t e = cudaStreamEndCapture(dv.main, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) { ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); return QCF_ERR_CUDA; }
        e = cudaGraphInstantiate(&dv.graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { dv.graph = nullptr; ctx->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return QCF_ERR_CUDA; }
        dv.graph_mode = gmode; dv.graph_pa = pa; dv.graph_pb = pb;
    }
    CK(cudaGraphLaunch(dv.graph, ms));
    ++*graph_launches;
    return QCF_OK;
}

int ensure_incremental_buffers(qcf_ctx* ctx, int nk) {
    const size_t nn = (size_t)ctx->N * ctx->N;
    for (auto& dv : ctx->dev) {
        CK(cudaSetDevice(dv.device));
        for (int k = 0; k < nk; ++k)
            if (!dv.Pprev[k]) { CK(cudaMalloc(&dv.Pprev[k], nn * sizeof(double))); CK(cudaMemset(dv.Pprev[k], 0, nn * sizeof(double))); }
    }
    qcf_device& d0 = ctx->dev[0];
    CK(cudaSetDevice(d0.device));
    for (int k = 0; k < 2; ++k)
        if (!d0.Gprev[k]) { CK(cudaMalloc(&d0.Gprev[k], nn * sizeof(double))); CK(cudaMemset(d0.Gprev[k], 0, nn * sizeof(double))); }
    return QCF_OK;
}

// dPa/dPb: unscaled densities on dev[0]; results into dG0/dG1 (dev[0]).  incremental: the scaled densities become
// P - P_prev and the result is ADDED to dG0/dG1.  Asynchronous on `user`.
int run_build_impl(qcf_ctx* ctx, int mode, const double* dPa, const double* dPb, double* dG0, double* dG1, cudaStream_t user,
                   bool incremental) {
    const double h0 = now_ms();
    const int N = ctx->N, nd = (int)ctx->dev.size();
    const size_t nn = (size_t)N * N;
    const int nk = (mode == 1) ? 2 : 1;
    qcf_device& d0 = ctx->dev[0];
    CK(cudaSetDevice(d0.device));
    CK(cudaEventRecord(ctx->ev_t0, user));
    int graph_launches = 0;
    if (nd > 1) CK(cudaEventRecord(ctx->ev_in, user));
    std::atomic<int> graph_launches_atomic{0};
    // one device's share: replicate the density (peer copy dev[0] -> dev[d] over NVLink on dev[d]'s stream), replay the graph
    auto enqueue_one = [&](int d, std::string& err) -> int {
        qcf_device& dv = ctx->dev[d];
        cudaStream_t sd = d == 0 ? user : dv.main;
        const double *pa = dPa, *pb = dPb;
        cudaError_t e;
#define CKD(call) do { e = (call); if (e != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e); return QCF_ERR_CUDA; } } while (0)
        CKD(cudaSetDevice(dv.device));
        if (d > 0) {
            CKD(cudaStreamWaitEvent(sd, ctx->ev_in, 0));
            CKD(cudaMemcpyPeerAsync(dv.Pin[0], dv.device, dPa, d0.device, nn * sizeof(double), sd));
            if (nk == 2) CKD(cudaMemcpyPeerAsync(dv.Pin[1], dv.device, dPb, d0.device, nn * sizeof(double), sd));
            pa = dv.Pin[0]; pb = nk == 2 ? dv.Pin[1] : nullptr;
        }
        CKD(cudaEventRecord(dv.ev_k0, sd));
        int gl = 0;
        int rc = launch_device_work(ctx, dv, mode, pa, pb, incremental, sd, &gl);
        if (rc) { err = ctx->err; return rc; }
        if (gl) ++graph_launches_atomic;
        CKD(cudaEventRecord(dv.ev_k1, sd));
        if (nd > 1) CKD(cudaEventRecord(dv.ev_acc, sd));
#undef CKD
        return QCF_OK;
    };
    // The devices are enqueued one after the other from the calling thread (~50 us of host time per graph launch).  Handing
    // them to the copy pool's threads was measured and is SLOWER (0.69-0.81 ms instead of 0.36-0.40 ms of host time for eight
    // devices: the launches serialise inside the driver and the hand-off adds wake-up latency), so it was not kept.
    for (int d = 0; d < nd; ++d) {
        std::string err;
        int rc = enqueue_one(d, err);
        if (rc) { ctx->err = err; return rc; }
    }
    graph_launches = graph_launches_atomic.load();
    PeerAcc acc{};
    for (int d = 0; d < nd; ++d) { acc.AJ[d] = ctx->dev[d].AJ; acc.AK0[d] = ctx->dev[d].AK[0]; acc.AK1[d] = ctx->dev[d].AK[1]; }
    const int tpb = 256;
    for (int d = 0; d < nd; ++d) {
        qcf_device& dv = ctx->dev[d];
        cudaStream_t sd = d == 0 ? user : dv.main;
        CK(cudaSetDevice(dv.device));
        for (int p = 0; p < nd; ++p)
            if (p != d) CK(cudaStreamWaitEvent(sd, ctx->dev[p].ev_acc, 0));
        const int row0 = (int)((long long)N * d / nd), row1 = (int)((long long)N * (d + 1) / nd);
        const size_t cnt = (size_t)(row1 - row0) * N;
        if (cnt) finalize_kernel<<<(int)((cnt + tpb - 1) / tpb), tpb, 0, sd>>>(N, mode, nd, acc, dv.fscale, dv.sc, row0, row1, dG0, dG1, incremental ? 1 : 0);
        CK(cudaGetLastError());
        if (nd > 1) CK(cudaEventRecord(dv.ev_red, sd));
    }
    CK(cudaSetDevice(d0.device));
    for (int d = 1; d < nd; ++d) CK(cudaStreamWaitEvent(user, ctx->dev[d].ev_red, 0));
    CK(cudaEventRecord(ctx->ev_t1, user));
    ctx->launches.clear();
    for (const PlannedLaunch& pl : ctx->plan) ctx->launches.push_back({pl.gi, pl.gj});
    int launches = 0;
    for (auto& dv : ctx->dev) launches += dv.launches + 1;
    ctx->stats.launches = launches;
    ctx->stats.graph_launches = graph_launches;
    ctx->stats.host_ms = now_ms() - h0;
    ctx->counters_pending = true;
    return QCF_OK;
}

int host_build(qcf_ctx* ctx, int mode, const double* Pa, const double* Pb, double* G0, double* G1, int incremental, int reset) {
    const size_t nn = (size_t)ctx->N * ctx->N;
    const int nk = (mode == 1) ? 2 : 1;
    qcf_device& d0 = ctx->dev[0];
    CK(cudaSetDevice(d0.device));
    cudaStream_t ms = d0.main;
    bool full_rebuild = false;
    if (incremental) {
        int rc = ensure_incremental_buffers(ctx, nk);
        if (rc) return rc;
        full_rebuild = reset || ctx->incremental_builds == 0;
    }
    CK(cudaEventRecord(ctx->ev_h0, ms));
    // Staging through pinned memory in pieces: the DMA of one piece runs while the host threads copy the next one
    // (matrices below 2 MB go in one piece).
    const int nchunk = nn * sizeof(double) >= (size_t(2) << 20) ? QCF_STAGE_CHUNKS : 1;
    auto piece = [&](int c) { return nn * (size_t)c / nchunk; };
    for (int k = 0; k < (Pb ? 2 : 1); ++k) {
        const double* src = k == 0 ? Pa : Pb;
        for (int c = 0; c < nchunk; ++c) {
            const size_t lo = piece(c), len = piece(c + 1) - lo;
            ctx->pool->copy(ctx->h_pin + k * nn + lo, src + lo, len * sizeof(double));
            CK(cudaMemcpyAsync(d0.Pin[k] + lo, ctx->h_pin + k * nn + lo, len * sizeof(double), cudaMemcpyHostToDevice, ms));
        }
    }
    double* g0 = incremental ? d0.Gprev[0] : d0.G[0];
    double* g1 = incremental ? d0.Gprev[1] : d0.G[1];
    int rc = incremental ? qcf_internal::run_build_scf(ctx, mode, d0.Pin[0], Pb ? d0.Pin[1] : nullptr, g0, g1, ms, true, full_rebuild)
                         : run_build_impl(ctx, mode, d0.Pin[0], Pb ? d0.Pin[1] : nullptr, g0, g1, ms, false);
    if (rc) return rc;
    if (incremental) ++ctx->incremental_builds;
    // ... and back the same way: every piece of G is copied out of the pinned buffer while the next one is in flight
    const int nout = G1 ? 2 : 1;
    for (int k = 0; k < nout; ++k)
        for (int c = 0; c < nchunk; ++c) {
            const size_t lo = piece(c), len = piece(c + 1) - lo;
            CK(cudaMemcpyAsync(ctx->h_pin + (2 + k) * nn + lo, (k == 0 ? g0 : g1) + lo, len * sizeof(double), cudaMemcpyDeviceToHost, ms));
            CK(cudaEventRecord(ctx->ev_out[k * QCF_STAGE_CHUNKS + c], ms));
        }
    CK(cudaEventRecord(ctx->ev_h1, ms));
    for (int k = 0; k < nout; ++k)
        for (int c = 0; c < nchunk; ++c) {
            const size_t lo = piece(c), len = piece(c + 1) - lo;
            CK(cudaEventSynchronize(ctx->ev_out[k * QCF_STAGE_CHUNKS + c]));
            ctx->pool->copy((k == 0 ? G0 : G1) + lo, ctx->h_pin + (2 + k) * nn + lo, len * sizeof(double));
        }
    CK(cudaStreamSynchronize(ms));
    float t = 0;
    CK(cudaEventElapsedTime(&t, ctx->ev_h0, ctx->ev_h1));
    ctx->stats.total_ms = t;
    return qcf_internal::collect_stats(ctx);
}

bool ready(qcf_ctx* ctx) {
    if (ctx->dev.empty() || !ctx->dev[0].AJ) { ctx->err = "context was not created successfully"; return false; }
    return true;
}

}  // namespace

namespace qcf_internal {

int run_build(qcf_ctx* ctx, int mode, const double* dPa, const double* dPb, double* dG0, double* dG1, cudaStream_t user) {
    return run_build_impl(ctx, mode, dPa, dPb, dG0, dG1, user, false);
}

int run_build_scf(qcf_ctx* ctx, int mode, const double* dPa, const double* dPb, double* dG0, double* dG1, cudaStream_t user,
                  bool incremental, bool reset) {
    if (!incremental) return run_build_impl(ctx, mode, dPa, dPb, dG0, dG1, user, false);
    const size_t nn = (size_t)ctx->N * ctx->N;
    const int nk = mode == 1 ? 2 : 1;
    int rc = ensure_incremental_buffers(ctx, nk);
    if (rc) return rc;
    if (reset) {
        // full rebuild: the plain build at tau into dG0/dG1, then P_prev <- P on every device (device 0 from the
        // caller's buffers, the others from the replicas the build has just peer-copied, each on the stream that
        // device's next build runs on)
        rc = run_build_impl(ctx, mode, dPa, dPb, dG0, dG1, user, false);
        if (rc) return rc;
        qcf_device& d0 = ctx->dev[0];
        CK(cudaSetDevice(d0.device));
        CK(cudaMemcpyAsync(d0.Pprev[0], dPa, nn * sizeof(double), cudaMemcpyDeviceToDevice, user));
        if (nk == 2) CK(cudaMemcpyAsync(d0.Pprev[1], dPb, nn * sizeof(double), cudaMemcpyDeviceToDevice, user));
        for (size_t d = 1; d < ctx->dev.size(); ++d) {
            qcf_device& dv = ctx->dev[d];
            CK(cudaSetDevice(dv.device));
            for (int k = 0; k < nk; ++k) CK(cudaMemcpyAsync(dv.Pprev[k], dv.Pin[k], nn * sizeof(double), cudaMemcpyDeviceToDevice, dv.main));
        }
        CK(cudaSetDevice(d0.device));
        return QCF_OK;
    }
    return run_build_impl(ctx, mode, dPa, dPb, dG0, dG1, user, true);
}

int collect_stats(qcf_ctx* ctx) {
    if (!ctx->counters_pending) return QCF_OK;
    qcf_device& d0 = ctx->dev[0];
    CK(cudaSetDevice(d0.device));
    CK(cudaEventSynchronize(ctx->ev_t1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
    ctx->stats.kernel_ms = ms;
    const int nl = (int)ctx->launches.size();
    std::vector<unsigned long long> cnt(std::max(nl, 1), 0ull), tmp(std::max(nl, 1));
    for (auto& dv : ctx->dev) {
        CK(cudaSetDevice(dv.device));
        CK(cudaEventSynchronize(dv.ev_k1));
        CK(cudaEventElapsedTime(&dv.last_ms, dv.ev_k0, dv.ev_k1));
        if (nl) CK(cudaMemcpy(tmp.data(), dv.counters, sizeof(unsigned long long) * nl, cudaMemcpyDeviceToHost));
        for (int i = 0; i < nl; ++i) cnt[i] += tmp[i];
    }
    CK(cudaSetDevice(d0.device));
    long long q = 0;
    double flops = 0;
    const double digest = 12.0;  // per unique contracted integral (RHF model)
    ctx->launch_cnt.assign(cnt.begin(), cnt.begin() + nl);
    for (int i = 0; i < nl; ++i) {
        if (ctx->profile && (int)d0.prof_ev.size() >= 2 * (i + 1)) {
            float t = 0;
            if (cudaEventElapsedTime(&t, d0.prof_ev[2 * i], d0.prof_ev[2 * i + 1]) == cudaSuccess) ctx->launches[i].ms = t;
            else cudaGetLastError();
        }
        const HostGroup& b = ctx->groups[ctx->launches[i].bra];
        const HostGroup& k = ctx->groups[ctx->launches[i].ket];
        q += (long long)cnt[i];
        const double nint = (double)ncart(b.la) * ncart(b.lb) * ncart(k.la) * ncart(k.lb);
        flops += (double)cnt[i] * ((double)b.K * k.K * model_flops_prim(b.la, b.lb, k.la, k.lb) + digest * nint);
    }
    ctx->stats.quartets = q;
    ctx->stats.model_flops = flops;
    ctx->counters_pending = false;
    return QCF_OK;
}

}  // namespace qcf_internal

// ======================================================================================================
extern "C" {

int qcf_create(const qcf_basis* b, const qcf_opts* o, qcf_ctx** out) {
    if (!b || !out) return QCF_ERR_ARG;
    *out = nullptr;
    const double t_create0 = now_ms();
    qcf_ctx* ctx = new qcf_ctx();
    auto fail = [&](int code, const std::string& msg) {
        // keep the context alive so that qcf_last_error can report; caller destroys it
        ctx->err = msg;
        *out = ctx;
        return code;
    };
    if (b->cartesian != 1) return fail(QCF_ERR_ARG, "only Cartesian (6d) shells are supported");
    if (b->n_shells <= 0 || b->n_atoms <= 0) return fail(QCF_ERR_ARG, "empty basis");
    if (!b->xyz || !b->shell_atom || !b->shell_l || !b->shell_nprim || !b->shell_prim_off || !b->exps || !b->coefs)
        return fail(QCF_ERR_ARG, "null array in qcf_basis");
    int device0 = 0, prank = 0, pworld = 1, ngpus = 1;
    if (o) {
        device0 = o->device; prank = o->rank; pworld = o->world_size > 0 ? o->world_size : 1;
        ngpus = o->n_gpus > 1 ? o->n_gpus : 1;
        ctx->deterministic = o->deterministic != 0;
        if (o->block_threads > 0) ctx->block = o->block_threads;
        if (o->screen_tau < -0.5) ctx->screening = false;
        else if (o->screen_tau > 0) ctx->tau = o->screen_tau;
    }
    if (const char* e = getenv("QCF_PROFILE")) ctx->profile = (e[0] == '1');
    if (const char* e = getenv("QCF_NO_GRAPH")) ctx->use_graph = !(e[0] == '1');
    if (const char* e = getenv("QCF_DETERMINISTIC")) ctx->deterministic = ctx->deterministic || (e[0] == '1');
    if (prank < 0 || prank >= pworld) return fail(QCF_ERR_ARG, "rank outside [0, world_size)");
    if (ngpus > QCF_MAXDEV) return fail(QCF_ERR_ARG, "n_gpus too large");
    if (device0 < 0) return fail(QCF_ERR_ARG, "negative device ordinal");
    ctx->world = pworld * ngpus;
    // (launch-shape defaults are chosen from the problem size after the pair lists exist, see tune_launch_shape below)
    if (const char* e = getenv("QCF_KETS_PER_THREAD")) ctx->kets_per_thread = std::max(1, atoi(e));
    if (const char* e = getenv("QCF_STREAMS")) ctx->nstreams = std::min(QCF_MAXSTREAM, std::max(1, atoi(e)));
    if (const char* e = getenv("QCF_SERIAL_CAP")) ctx->serial_cap = std::max(1.0, atof(e));
    if (const char* e = getenv("QCF_TARGET_CTAS")) ctx->target_ctas = std::max(1, atoi(e));
    if (const char* e = getenv("QCF_PS_MIN")) ctx->ps_min_prim = std::max(1, atoi(e));
    if (const char* e = getenv("QCF_BLOCK")) ctx->block = atoi(e);
    if (const char* e = getenv("QCF_SPLIT_MIN_BRAS")) ctx->split_min_bras = std::max(0, atoi(e));
    if (const char* e = getenv("QCF_RED_EPS_FACTOR")) ctx->red_eps_factor = std::max(0.0, atof(e));
    if (const char* e = getenv("QCF_PRIM_CUT")) ctx->prim_cut_factor = std::max(0.0, atof(e));
    if (const char* e = getenv("QCF_PAIR_CUT")) ctx->pair_cut_factor = std::max(0.0, atof(e));
    if (const char* e = getenv("QCF_ORDER")) ctx->launch_order = atoi(e);
    if (const char* e = getenv("QCF_PRIO")) ctx->stream_prio = atoi(e);
    if (ctx->block < 32 || ctx->block > 128 || ctx->block % 32) return fail(QCF_ERR_ARG, "block_threads must be 32, 64, 96 or 128");
    ctx->natoms = b->n_atoms; ctx->nshell = b->n_shells;
    ctx->xyz.assign(b->xyz, b->xyz + 3 * b->n_atoms);
    ctx->charge.resize(b->n_atoms);
    for (int i = 0; i < b->n_atoms; ++i) ctx->charge[i] = b->Z ? (double)b->Z[i] : 0.0;
    ctx->sh_atom.assign(b->shell_atom, b->shell_atom + b->n_shells);
    ctx->sh_l.assign(b->shell_l, b->shell_l + b->n_shells);
    ctx->sh_np.assign(b->shell_nprim, b->shell_nprim + b->n_shells);
    ctx->sh_po.assign(b->shell_prim_off, b->shell_prim_off + b->n_shells);
    long long nprim = 0;
    ctx->sh_off.assign(b->n_shells + 1, 0);
    for (int s = 0; s < b->n_shells; ++s) {
        if (ctx->sh_l[s] < 0 || ctx->sh_l[s] > LMAX) return fail(QCF_ERR_ARG, "angular momentum outside 0..2");
        if (ctx->sh_np[s] <= 0) return fail(QCF_ERR_ARG, "shell without primitives");
        // primitive-pair indices are stored as uint16 (HostPair::order): K = np_a * np_b must fit
        if (ctx->sh_np[s] > 255) return fail(QCF_ERR_ARG, "more than 255 primitives in one shell");
        if (ctx->sh_po[s] < 0) return fail(QCF_ERR_ARG, "negative shell_prim_off");
        if (ctx->sh_atom[s] < 0 || ctx->sh_atom[s] >= b->n_atoms) return fail(QCF_ERR_ARG, "shell_atom out of range");
        nprim = std::max(nprim, (long long)ctx->sh_po[s] + ctx->sh_np[s]);
        ctx->sh_off[s + 1] = ctx->sh_off[s] + ncart(ctx->sh_l[s]);
    }
    if (nprim > (1ll << 30)) return fail(QCF_ERR_ARG, "primitive offsets out of range");
    ctx->exps.assign(b->exps, b->exps + nprim);
    ctx->coefs.assign(b->coefs, b->coefs + nprim);
    for (long long k = 0; k < nprim; ++k)
        if (!(ctx->exps[k] > 0.0) || !std::isfinite(ctx->exps[k]) || !std::isfinite(ctx->coefs[k]))
            return fail(QCF_ERR_ARG, "exponents must be positive and finite, coefficients finite");
    ctx->N = ctx->sh_off.back();
    ctx->fscale.resize(ctx->N);
    for (int s = 0; s < b->n_shells; ++s) {
        const int l = ctx->sh_l[s];
        for (int i = 0; i < ncart(l); ++i) {
            // N(i,j,k)/N(l,0,0) = sqrt((2l-1)!! / ((2i-1)!!(2j-1)!!(2k-1)!!)); for l<=2 only xy,xz,yz differ: sqrt(3)
            const int x = cart_x(l, i), y = cart_y(l, i), z = cart_z(l, i);
            ctx->fscale[ctx->sh_off[s] + i] = (l == 2 && x < 2 && y < 2 && z < 2) ? std::sqrt(3.0) : 1.0;
        }
    }
    *out = ctx;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        ctx->err = "no CUDA device available (this engine has no CPU fallback)";
        return QCF_ERR_CUDA;
    }
    if (device0 + ngpus > ndev) {
        ctx->err = "n_gpus exceeds the CUDA devices visible to this process";
        return QCF_ERR_ARG;
    }
    ctx->dev.resize(ngpus);
    for (int d = 0; d < ngpus; ++d) { ctx->dev[d].device = device0 + d; ctx->dev[d].rank = prank * ngpus + d; }
    CK(cudaSetDevice(device0));
    const size_t nn = (size_t)ctx->N * ctx->N;
    std::vector<double> tab = make_boys_table();
    CK(upload(&ctx->dev[0].boys, tab));
    CK(cudaEventCreate(&ctx->ev_t0)); CK(cudaEventCreate(&ctx->ev_t1));
    CK(cudaEventCreate(&ctx->ev_h0)); CK(cudaEventCreate(&ctx->ev_h1));
    CK(cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming));
    for (cudaEvent_t& e : ctx->ev_out) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int rc = build_pairs(ctx);
    if (rc) return rc;
    tune_launch_shape(ctx);
    make_plan(ctx);
    // peer access between all devices of the context (NVLink / NVSwitch): the finalize kernel reads every device's
    // accumulators directly
    for (int d = 0; d < ngpus; ++d)
        for (int p = 0; p < ngpus; ++p) {
            if (p == d) continue;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, ctx->dev[d].device, ctx->dev[p].device));
            if (!can) { ctx->err = "n_gpus > 1 needs peer access between the devices (NVLink)"; return QCF_ERR_CUDA; }
            CK(cudaSetDevice(ctx->dev[d].device));
            cudaError_t e = cudaDeviceEnablePeerAccess(ctx->dev[p].device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return QCF_ERR_CUDA; }
        }
    for (int d = 0; d < ngpus; ++d) {
        rc = setup_device(ctx, ctx->dev[d], tab);
        if (rc) return rc;
    }
    CK(cudaSetDevice(device0));
    CK(cudaHostAlloc(&ctx->h_pin, 4 * nn * sizeof(double), cudaHostAllocPortable));
    {
        int nthr = 3;     // helper threads of the staging copies (plus the calling thread)
        if (const char* e = getenv("QCF_COPY_THREADS")) nthr = std::max(0, std::min(15, atoi(e) - 1));
        ctx->pool = new qcf_copy_pool(nthr);
    }
    ctx->stats.n_basis = ctx->N; ctx->stats.n_shells = ctx->nshell;
    ctx->stats.n_devices = ngpus;
    const long long nsp = (long long)ctx->nshell * (ctx->nshell + 1) / 2;
    ctx->stats.quartets_total = nsp * (nsp + 1) / 2;
    double cmax = 0, csum = 0;
    for (double c : ctx->rank_cost) { cmax = std::max(cmax, c); csum += c; }
    ctx->stats.rank_imbalance = csum > 0 ? cmax / (csum / ctx->rank_cost.size()) : 1.0;
    for (auto& dv : ctx->dev) { CK(cudaSetDevice(dv.device)); CK(cudaDeviceSynchronize()); }
    CK(cudaSetDevice(device0));
    ctx->stats.create_ms = now_ms() - t_create0;
    return QCF_OK;
}

int qcf_nbasis(const qcf_ctx* ctx) { return ctx ? ctx->N : QCF_ERR_ARG; }

int qcf_build_rhf(qcf_ctx* ctx, const double* P, double* G) {
    if (!ctx || !P || !G) return QCF_ERR_ARG;
    if (!ready(ctx)) return QCF_ERR_STATE;
    return host_build(ctx, 0, P, nullptr, G, nullptr, 0, 0);
}

int qcf_build_uhf(qcf_ctx* ctx, const double* Pa, const double* Pb, double* Ga, double* Gb) {
    if (!ctx || !Pa || !Pb || !Ga || !Gb) return QCF_ERR_ARG;
    if (!ready(ctx)) return QCF_ERR_STATE;
    return host_build(ctx, 1, Pa, Pb, Ga, Gb, 0, 0);
}

int qcf_build_rhf_incremental(qcf_ctx* ctx, const double* P, double* G, int reset) {
    if (!ctx || !P || !G) return QCF_ERR_ARG;
    if (!ready(ctx)) return QCF_ERR_STATE;
    return host_build(ctx, 0, P, nullptr, G, nullptr, 1, reset);
}

int qcf_build_uhf_incremental(qcf_ctx* ctx, const double* Pa, const double* Pb, double* Ga, double* Gb, int reset) {
    if (!ctx || !Pa || !Pb || !Ga || !Gb) return QCF_ERR_ARG;
    if (!ready(ctx)) return QCF_ERR_STATE;
    return host_build(ctx, 1, Pa, Pb, Ga, Gb, 1, reset);
}

int qcf_build_jk(qcf_ctx* ctx, int nd, const double* const* P, double* const* J, double* const* K) {
    if (!ctx || nd <= 0 || !P || !J || !K) return QCF_ERR_ARG;
    if (!ready(ctx)) return QCF_ERR_STATE;
    for (int d = 0; d < nd; ++d) {
        if (!P[d] || !J[d] || !K[d]) return QCF_ERR_ARG;
        int rc = host_build(ctx, 2, P[d], nullptr, J[d], K[d], 0, 0);
        if (rc) return rc;
    }
    return QCF_OK;
}

int qcf_build_rhf_dev(qcf_ctx* ctx, const double* dP, double* dG, void* stream) {
    if (!ctx || !dP || !dG) return QCF_ERR_ARG;
    if (!ready(ctx)) return QCF_ERR_STATE;
    return qcf_internal::run_build(ctx, 0, dP, nullptr, dG, nullptr, (cudaStream_t)stream);
}

int qcf_build_uhf_dev(qcf_ctx* ctx, const double* dPa, const double* dPb, double* dGa, double* dGb, void* stream) {
    if (!ctx || !dPa || !dPb || !dGa || !dGb) return QCF_ERR_ARG;
    if (!ready(ctx)) return QCF_ERR_STATE;
    return qcf_internal::run_build(ctx, 1, dPa, dPb, dGa, dGb, (cudaStream_t)stream);
}

int qcf_eri_quartet(qcf_ctx* ctx, int s1, int s2, int s3, int s4, double* out) {
    if (!ctx || !out) return QCF_ERR_ARG;
    if (!ready(ctx)) return QCF_ERR_STATE;
    const int ns = ctx->nshell;
    if (s1 < 0 || s2 < 0 || s3 < 0 || s4 < 0 || s1 >= ns || s2 >= ns || s3 >= ns || s4 >= ns) return QCF_ERR_ARG;
    qcf_device& d0 = ctx->dev[0];
    CK(cudaSetDevice(d0.device));
    // canonical pairs: higher l first, ties by larger shell index first (as build_pairs does)
    auto canon = [&](int x, int y, bool& swapped) {
        int a = std::max(x, y), b = std::min(x, y);
        if (ctx->sh_l[b] > ctx->sh_l[a]) std::swap(a, b);
        swapped = (a != x);
        return std::make_pair(a, b);
    };
    bool sw12, sw34;
    auto p12 = canon(s1, s2, sw12), p34 = canon(s3, s4, sw34);
    const int n1 = ncart(ctx->sh_l[s1]), n2 = ncart(ctx->sh_l[s2]), n3 = ncart(ctx->sh_l[s3]), n4 = ncart(ctx->sh_l[s4]);
    const size_t ntot = (size_t)n1 * n2 * n3 * n4;
    auto i12 = ctx->pair_index.find(p12), i34 = ctx->pair_index.find(p34);
    if (i12 == ctx->pair_index.end() || i34 == ctx->pair_index.end()) {   // screened-out pair
        std::fill(out, out + ntot, 0.0);
        return QCF_OK;
    }
    const HostGroup& g12 = ctx->groups[i12->second.first];
    const HostGroup& g34 = ctx->groups[i34->second.first];
    const bool bra12 = g12.cls >= g34.cls;
    const int gb = bra12 ? i12->second.first : i34->second.first;
    const int gk = bra12 ? i34->second.first : i12->second.first;
    const int ib = bra12 ? i12->second.second : i34->second.second;
    const int ik = bra12 ? i34->second.second : i12->second.second;
    double* dout = nullptr;
    CK(cudaMalloc(&dout, ntot * sizeof(double)));
    class_table(ctx->groups[gb].cls, ctx->groups[gk].cls)->quartet(0, d0.groups[gb].pg, ib, d0.groups[gk].pg, ik, d0.boys, dout);
    CK(cudaGetLastError());
    std::vector<double> tmp(ntot);
    CK(cudaMemcpy(tmp.data(), dout, ntot * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dout);
    // kernel layout: [bra a][bra b][ket c][ket d] with (a,b),(c,d) canonical; map back to (s1 s2|s3 s4)
    const int o1 = ctx->sh_off[s1], o2 = ctx->sh_off[s2], o3 = ctx->sh_off[s3], o4 = ctx->sh_off[s4];
    for (int i = 0; i < n1; ++i) for (int j = 0; j < n2; ++j) for (int k = 0; k < n3; ++k) for (int l = 0; l < n4; ++l) {
        // indices in canonical pair order
        const int a12 = sw12 ? j : i, b12 = sw12 ? i : j, nb12 = sw12 ? n1 : n2;
        const int a34 = sw34 ? l : k, b34 = sw34 ? k : l, nb34 = sw34 ? n3 : n4;
        size_t idx;
        if (bra12) idx = (((size_t)a12 * nb12 + b12) * (sw34 ? n4 : n3) + a34) * nb34 + b34;
        else idx = (((size_t)a34 * nb34 + b34) * (sw12 ? n2 : n1) + a12) * nb12 + b12;
        const double s = ctx->fscale[o1 + i] * ctx->fscale[o2 + j] * ctx->fscale[o3 + k] * ctx->fscale[o4 + l];
        out[(((size_t)i * n2 + j) * n3 + k) * n4 + l] = s * tmp[idx];
    }
    return QCF_OK;
}

int qcf_one_electron(qcf_ctx* ctx, double* S, double* T, double* V) {
    if (!ctx || !S || !T || !V) return QCF_ERR_ARG;
    if (!ready(ctx)) return QCF_ERR_STATE;
    qcf_device& d0 = ctx->dev[0];
    CK(cudaSetDevice(d0.device));
    const size_t nn = (size_t)ctx->N * ctx->N;
    int *d_i = nullptr;
    double* d_d = nullptr;
    const int ns = ctx->nshell, na = ctx->natoms;
    std::vector<int> hi;
    hi.insert(hi.end(), ctx->sh_atom.begin(), ctx->sh_atom.end());
    hi.insert(hi.end(), ctx->sh_l.begin(), ctx->sh_l.end());
    hi.insert(hi.end(), ctx->sh_np.begin(), ctx->sh_np.end());
    hi.insert(hi.end(), ctx->sh_po.begin(), ctx->sh_po.end());
    std::vector<double> hd;
    hd.insert(hd.end(), ctx->exps.begin(), ctx->exps.end());
    hd.insert(hd.end(), ctx->coefs.begin(), ctx->coefs.end());
    hd.insert(hd.end(), ctx->xyz.begin(), ctx->xyz.end());
    hd.insert(hd.end(), ctx->charge.begin(), ctx->charge.end());
    CK(upload(&d_i, hi));
    CK(upload(&d_d, hd));
    double* d_out = nullptr;
    CK(cudaMalloc(&d_out, 3 * nn * sizeof(double)));
    CK(cudaMemset(d_out, 0, 3 * nn * sizeof(double)));
    ShellData sd{};
    sd.nshell = ns; sd.natoms = na; sd.N = ctx->N;
    sd.atom = d_i; sd.l = d_i + ns; sd.nprim = d_i + 2 * ns; sd.prim_off = d_i + 3 * ns; sd.off = d0.shoff;
    const size_t npr = ctx->exps.size();
    sd.exps = d_d; sd.coefs = d_d + npr; sd.xyz = d_d + 2 * npr; sd.charge = d_d + 2 * npr + 3 * (size_t)na;
    sd.fscale = d0.fscale;
    const long long npair = (long long)ns * (ns + 1) / 2;
    const long long nblk = (npair * 32 + 127) / 128;
    onee_kernel<<<(unsigned)nblk, 128>>>(sd, d0.boys, d_out, d_out + nn, d_out + 2 * nn);
    CK(cudaGetLastError());
    CK(cudaMemcpy(S, d_out, nn * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(T, d_out + nn, nn * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(V, d_out + 2 * nn, nn * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_out); cudaFree(d_i); cudaFree(d_d);
    return QCF_OK;
}

int qcf_schwarz(qcf_ctx* ctx, double* Q) {
    if (!ctx || !Q) return QCF_ERR_ARG;
    const int ns = ctx->nshell;
    std::fill(Q, Q + (size_t)ns * ns, 0.0);
    for (const auto& g : ctx->groups)
        for (const auto& p : g.pairs) { Q[(size_t)p.sa * ns + p.sb] = p.Q; Q[(size_t)p.sb * ns + p.sa] = p.Q; }
    return QCF_OK;
}

int qcf_boys(qcf_ctx* ctx, int mmax, int n, const double* T, double* F) {
    if (!ctx || !T || !F || mmax < 0 || mmax > BOYS_LTOT || n <= 0) return QCF_ERR_ARG;
    if (!ready(ctx)) return QCF_ERR_STATE;
    qcf_device& d0 = ctx->dev[0];
    CK(cudaSetDevice(d0.device));
    double *dT = nullptr, *dF = nullptr;
    CK(cudaMalloc(&dT, sizeof(double) * n));
    CK(cudaMalloc(&dF, sizeof(double) * n * (mmax + 1)));
    CK(cudaMemcpy(dT, T, sizeof(double) * n, cudaMemcpyHostToDevice));
    const int g = (n + 127) / 128;
    switch (mmax) {
#define QCF_B(L) case L: boys_test_kernel<L><<<g, 128>>>(n, dT, d0.boys, dF); break;
        QCF_B(0) QCF_B(1) QCF_B(2) QCF_B(3) QCF_B(4) QCF_B(5) QCF_B(6) QCF_B(7) QCF_B(8)
    }
    CK(cudaGetLastError());
    CK(cudaMemcpy(F, dF, sizeof(double) * n * (mmax + 1), cudaMemcpyDeviceToHost));
    cudaFree(dT); cudaFree(dF);
    return QCF_OK;
}

int qcf_fp64_peak(qcf_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return QCF_ERR_ARG;
    if (!ready(ctx)) return QCF_ERR_STATE;
    const int device = ctx->dev[0].device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double* dout = nullptr;
    CK(cudaMalloc(&dout, sizeof(double)));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0, 0));
        fp64_peak_kernel<<<blocks, threads>>>(dout, iters, 1.0000001);
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(dout);
    *tflops = best;
    return QCF_OK;
}

int qcf_stats(const qcf_ctx* cctx, qcf_stats_t* out) {
    if (!cctx || !out) return QCF_ERR_ARG;
    qcf_ctx* ctx = const_cast<qcf_ctx*>(cctx);
    int rc = qcf_internal::collect_stats(ctx);
    if (rc) return rc;
    *out = ctx->stats;
    return QCF_OK;
}

int qcf_device_times(qcf_ctx* ctx, int max_dev, double* ms) {
    if (!ctx || (max_dev > 0 && !ms)) return QCF_ERR_ARG;
    int rc = qcf_internal::collect_stats(ctx);
    if (rc) return rc;
    const int nd = (int)ctx->dev.size();
    for (int d = 0; d < nd && d < max_dev; ++d) ms[d] = ctx->dev[d].last_ms;
    return nd;
}

int qcf_launch_profile(qcf_ctx* ctx, int max_rec, qcf_launch_rec* out) {
    if (!ctx || (max_rec > 0 && !out)) return QCF_ERR_ARG;
    int rc = qcf_internal::collect_stats(ctx);
    if (rc) return rc;
    const int nl = (int)ctx->launches.size();
    for (int i = 0; i < nl && i < max_rec; ++i) {
        const HostGroup& b = ctx->groups[ctx->launches[i].bra];
        const HostGroup& k = ctx->groups[ctx->launches[i].ket];
        out[i] = {b.la, b.lb, b.K, k.la, k.lb, k.K, (int)b.pairs.size(), (int)k.pairs.size(),
                  i < (int)ctx->launch_cnt.size() ? (long long)ctx->launch_cnt[i] : 0,
                  model_flops_prim(b.la, b.lb, k.la, k.lb), ctx->launches[i].ms};
    }
    return nl;
}

const char* qcf_last_error(const qcf_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void qcf_destroy(qcf_ctx* ctx) {
    if (!ctx) return;
    qcf_internal::scf_destroy(ctx);
    for (auto& dv : ctx->dev) destroy_device(dv);
    if (!ctx->dev.empty()) cudaSetDevice(ctx->dev[0].device);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    delete ctx->pool;
    for (cudaEvent_t e : {ctx->ev_t0, ctx->ev_t1, ctx->ev_h0, ctx->ev_h1, ctx->ev_in}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_out) if (e) cudaEventDestroy(e);
    delete ctx;
}

}  // extern "C"
