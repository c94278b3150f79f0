// scf_device.cu -- the per-iteration dense linear algebra of the reference's SCF loops on the device.
//
// With the Fock build at tens of milliseconds, the host eigensolve and the five N x N products of every iteration
// (core/src/hf/rhf.rs:70-88, uhf.rs:81-135, utils.rs:15-36) and the two PCIe crossings of P and G become the
// bottleneck.  qcf_scf_init / qcf_scf_step keep P, G, F, the DIIS history and the orbitals in HBM: per iteration only
// a handful of scalars (DIIS dot products, energy, density change) cross to the host.
//
// This is dense FP64 linear algebra, not the hot kernel: products are cuBLAS DGEMM, the symmetric eigensolve is
// cuSOLVER DSYEVD (the reference uses nalgebra's SymmetricEigen + a sort, utils.rs:20-36; DSYEVD returns ascending
// eigenvalues).  Every parity-relevant quirk of the reference loop is kept (list in qchem-rs_b200/hf.py):
//   Hueckel guess with the diagonal scaled by 1.75 too (rhf.rs:139-143), DIIS(4,6) RHF / DIIS(2,8) per spin UHF
//   (rhf.rs:65, uhf.rs:76-78), newest-first samples, +1 border, QR solve (diis.rs:29-51), energy from the NEW density
//   and the OLD G (rhf.rs:84-85), diagonal-only density rms (rhf.rs:87-88), UHF rms halved twice (uhf.rs:137,139).
#include "engine_internal.h"

#include <cublas_v2.h>
#include <cusolverDn.h>

#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <deque>

namespace {

struct Sample { double* err; double* fock; int id; };

struct Diis {
    int min_len = 4, max_len = 6;
    std::deque<Sample> samples;                 // newest first (diis.rs:29)
    std::vector<Sample> pool;                   // max_len + 1 preallocated slots
    std::map<std::pair<int, int>, double> dots; // cached <e_i, e_j> by sample id
    int next_id = 0;
};

}  // namespace

struct qcf_scf {
    int N = 0, nspin = 1, nocc[2] = {0, 0}, iteration = 0, full_every = 0, since_full = 0;
    double last_rms = 1.0;           // density rms of the previous step: difference-density builds pay off only once it is small
    double inc_rms = 1e-4;           // ... i.e. below this rms (QCF_INC_RMS)
    double factor = 2.0;
    cublasHandle_t blas = nullptr;
    cusolverDnHandle_t sol = nullptr;
    cudaStream_t stream = nullptr;
    double *S = nullptr, *H = nullptr, *X = nullptr, *T1 = nullptr, *T2 = nullptr, *Fx = nullptr, *C = nullptr, *W = nullptr;
    double *P[2] = {nullptr, nullptr}, *G[2] = {nullptr, nullptr}, *F[2] = {nullptr, nullptr}, *Pnext[2] = {nullptr, nullptr};
    double* work = nullptr;
    int lwork = 0;
    int* info = nullptr;
    double* scal = nullptr;          // device scalars: [0..15] dot products / sums
    Diis diis[2];
    std::vector<double> eps[2];      // orbital energies (host copy)
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
};

namespace {

#define CKB(call)                                                                                   \
    do {                                                                                            \
        cublasStatus_t s__ = (call);                                                                \
        if (s__ != CUBLAS_STATUS_SUCCESS) { ctx->err = std::string(#call) + ": cuBLAS status " + std::to_string((int)s__); return QCF_ERR_CUDA; } \
    } while (0)
#define CKS(call)                                                                                   \
    do {                                                                                            \
        cusolverStatus_t s__ = (call);                                                              \
        if (s__ != CUSOLVER_STATUS_SUCCESS) { ctx->err = std::string(#call) + ": cuSOLVER status " + std::to_string((int)s__); return QCF_ERR_CUDA; } \
    } while (0)
#define CK QCF_CK

__global__ void scale_columns_kernel(int N, double* __restrict__ U, const double* __restrict__ w) {
    // column-major U: U[:, j] *= 1 / sqrt(w[j])
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)N * N) return;
    const int j = idx / N;
    U[idx] *= 1.0 / sqrt(w[j]);
}

// h_eht_ij = 1.75 S_ij (H_ii + H_jj) / 2, diagonal included (rhf.rs:139-143)
__global__ void hueckel_kernel(int N, const double* __restrict__ S, const double* __restrict__ H, double* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)N * N) return;
    const int i = idx / N, j = idx % N;
    out[idx] = 1.75 * S[idx] * (H[(size_t)i * N + i] + H[(size_t)j * N + j]) / 2.0;
}

__global__ void add_kernel(size_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) out[idx] = a[idx] + b[idx];
}

// out = sum_i c[i] M_i (DIIS extrapolation, diis.rs:52-58), coefficients by value
struct Combo { const double* m[8]; double c[8]; int n; };
__global__ void combine_kernel(size_t nn, Combo cb, double* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nn) return;
    double s = 0.0;
    for (int i = 0; i < cb.n; ++i) s += cb.c[i] * cb.m[i][idx];
    out[idx] = s;
}

// sum_i (Pnew_ii - P_ii)^2, one block, fixed order
__global__ void diag_change_kernel(int N, const double* __restrict__ Pnew, const double* __restrict__ P, double* __restrict__ out) {
    __shared__ double part[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < N; i += 256) { const double d = Pnew[(size_t)i * N + i] - P[(size_t)i * N + i]; s += d * d; }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) *out = part[0];
}

inline int nblk(size_t n) { return (int)((n + 255) / 256); }

// C(col-major) = op(A) op(B)
int gemm(qcf_ctx* ctx, qcf_scf* s, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, double alpha, const double* A,
         const double* B, double beta, double* C) {
    const int N = s->N;
    CKB(cublasDgemm(s->blas, ta, tb, m, n, k, &alpha, A, N, B, N, &beta, C, N));
    return QCF_OK;
}

// eigenvectors (columns, ascending eigenvalues) of the symmetric matrix in A, in place; eigenvalues to s->W
int eigh(qcf_ctx* ctx, qcf_scf* s, double* A) {
    CKS(cusolverDnDsyevd(s->sol, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, s->N, A, s->N, s->W, s->work, s->lwork, s->info));
    return QCF_OK;
}

// P = factor * C[:, :nocc] C[:, :nocc]^T  (rhf.rs:169-181, uhf.rs:229-241)
int density_from(qcf_ctx* ctx, qcf_scf* s, const double* C, int nocc, double* P) {
    return gemm(ctx, s, CUBLAS_OP_N, CUBLAS_OP_T, s->N, s->N, nocc, s->factor, C, C, 0.0, P);
}

// C = X * eigvecs(X^T M X); M is overwritten (T1, T2 scratch)
int orbitals_of(qcf_ctx* ctx, qcf_scf* s, const double* M, double* Cout) {
    int rc = gemm(ctx, s, CUBLAS_OP_T, CUBLAS_OP_N, s->N, s->N, s->N, 1.0, s->X, M, 0.0, s->T1);
    if (rc) return rc;
    rc = gemm(ctx, s, CUBLAS_OP_N, CUBLAS_OP_N, s->N, s->N, s->N, 1.0, s->T1, s->X, 0.0, s->T2);
    if (rc) return rc;
    rc = eigh(ctx, s, s->T2);
    if (rc) return rc;
    return gemm(ctx, s, CUBLAS_OP_N, CUBLAS_OP_N, s->N, s->N, s->N, 1.0, s->X, s->T2, 0.0, Cout);
}

// Householder QR solve of the small DIIS system (diis.rs:48-51: matrix.qr().solve(&b)); false where nalgebra returns
// None (a zero on the diagonal of R)
bool qr_solve(int n, std::vector<double> A /* row-major n x n */, std::vector<double> b, std::vector<double>& x) {
    for (int k = 0; k < n; ++k) {
        double norm = 0;
        for (int i = k; i < n; ++i) norm += A[i * n + k] * A[i * n + k];
        norm = std::sqrt(norm);
        if (norm == 0.0) return false;
        const double alpha = A[k * n + k] > 0 ? -norm : norm;
        std::vector<double> v(n, 0.0);
        for (int i = k; i < n; ++i) v[i] = A[i * n + k];
        v[k] -= alpha;
        double vnorm2 = 0;
        for (int i = k; i < n; ++i) vnorm2 += v[i] * v[i];
        if (vnorm2 > 0) {
            for (int j = k; j < n; ++j) {
                double d = 0;
                for (int i = k; i < n; ++i) d += v[i] * A[i * n + j];
                d *= 2.0 / vnorm2;
                for (int i = k; i < n; ++i) A[i * n + j] -= d * v[i];
            }
            double d = 0;
            for (int i = k; i < n; ++i) d += v[i] * b[i];
            d *= 2.0 / vnorm2;
            for (int i = k; i < n; ++i) b[i] -= d * v[i];
        }
    }
    x.assign(n, 0.0);
    for (int i = n - 1; i >= 0; --i) {
        if (A[i * n + i] == 0.0) return false;
        double sum = b[i];
        for (int j = i + 1; j < n; ++j) sum -= A[i * n + j] * x[j];
        x[i] = sum / A[i * n + i];
    }
    return true;
}

// diis.rs:28-59 with the samples kept in HBM: pushes (err, fock) copies, returns the extrapolated Fock matrix in `out`
int diis_fock(qcf_ctx* ctx, qcf_scf* s, Diis& d, const double* err, const double* fock, double* out) {
    const size_t nn = (size_t)s->N * s->N;
    // take a free slot (the oldest sample when the ring is full, diis.rs:30 truncate)
    Sample slot;
    if ((int)d.samples.size() >= d.max_len) { slot = d.samples.back(); d.samples.pop_back(); }
    else slot = d.pool[d.samples.size()];
    slot.id = d.next_id++;
    CK(cudaMemcpyAsync(slot.err, err, nn * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    CK(cudaMemcpyAsync(slot.fock, fock, nn * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    d.samples.push_front(slot);
    const int n = (int)d.samples.size();
    // dot products <e_0, e_j> of the newest error matrix with every kept one (older pairs are cached); computed on every
    // call, also while n < min_len, because those samples enter the B matrix later
    CKB(cublasSetPointerMode(s->blas, CUBLAS_POINTER_MODE_DEVICE));
    for (int j = 0; j < n; ++j)
        CKB(cublasDdot(s->blas, (int)nn, d.samples[0].err, 1, d.samples[j].err, 1, s->scal + j));
    CKB(cublasSetPointerMode(s->blas, CUBLAS_POINTER_MODE_HOST));
    double hd[8];
    CK(cudaMemcpyAsync(hd, s->scal, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    for (int j = 0; j < n; ++j) {
        d.dots[{d.samples[0].id, d.samples[j].id}] = hd[j];
        d.dots[{d.samples[j].id, d.samples[0].id}] = hd[j];
    }
    if (n < d.min_len) {
        CK(cudaMemcpyAsync(out, fock, nn * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        return QCF_OK;
    }
    const int m = n + 1;
    std::vector<double> B((size_t)m * m, 0.0), rhs(m, 0.0), c;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            if (i == n && j == n) B[i * m + j] = 0.0;
            else if (i == n || j == n) B[i * m + j] = 1.0;
            else B[i * m + j] = d.dots[{d.samples[i].id, d.samples[j].id}];
        }
    rhs[n] = 1.0;
    if (!qr_solve(m, B, rhs, c)) { ctx->err = "DIIS failed (singular B matrix; rhs.rs:73 would panic)"; return QCF_ERR_STATE; }
    Combo cb{};
    cb.n = n;
    for (int i = 0; i < n; ++i) { cb.m[i] = d.samples[i].fock; cb.c[i] = c[i]; }
    combine_kernel<<<nblk(nn), 256, 0, s->stream>>>(nn, cb, out);
    CK(cudaGetLastError());
    // drop cached dots of evicted samples
    if (d.dots.size() > 400) {
        std::map<std::pair<int, int>, double> keep;
        for (auto& a : d.samples) for (auto& b2 : d.samples) { auto it = d.dots.find({a.id, b2.id}); if (it != d.dots.end()) keep.insert(*it); }
        d.dots.swap(keep);
    }
    return QCF_OK;
}

void free_scf(qcf_scf* s) {
    if (!s) return;
    for (double* p : {s->S, s->H, s->X, s->T1, s->T2, s->Fx, s->C, s->W, s->work, s->scal}) cudaFree(p);
    for (int k = 0; k < 2; ++k) {
        cudaFree(s->P[k]); cudaFree(s->G[k]); cudaFree(s->F[k]); cudaFree(s->Pnext[k]);
        for (auto& sm : s->diis[k].pool) { cudaFree(sm.err); cudaFree(sm.fock); }
    }
    cudaFree(s->info);
    if (s->blas) cublasDestroy(s->blas);
    if (s->sol) cusolverDnDestroy(s->sol);
    if (s->stream) cudaStreamDestroy(s->stream);
    for (cudaEvent_t e : {s->e0, s->e1, s->e2}) if (e) cudaEventDestroy(e);
    delete s;
}

}  // namespace

namespace qcf_internal {
void scf_destroy(qcf_ctx* ctx) {
    if (ctx && ctx->scf) {
        if (!ctx->dev.empty()) cudaSetDevice(ctx->dev[0].device);
        free_scf(ctx->scf);
        ctx->scf = nullptr;
    }
}
}  // namespace qcf_internal

extern "C" {

int qcf_scf_init(qcf_ctx* ctx, const double* S, const double* H, int unrestricted, int n_alpha, int n_beta, int full_rebuild_every) {
    if (!ctx || !S || !H) return QCF_ERR_ARG;
    if (ctx->dev.empty() || !ctx->dev[0].AJ) { ctx->err = "context was not created successfully"; return QCF_ERR_STATE; }
    const int N = ctx->N;
    if (n_alpha < 0 || n_alpha > N || n_beta < 0 || n_beta > N) return QCF_ERR_ARG;
    qcf_internal::scf_destroy(ctx);
    CK(cudaSetDevice(ctx->dev[0].device));
    qcf_scf* s = new qcf_scf();
    ctx->scf = s;
    s->N = N; s->nspin = unrestricted ? 2 : 1; s->nocc[0] = n_alpha; s->nocc[1] = unrestricted ? n_beta : n_alpha;
    s->factor = unrestricted ? 1.0 : 2.0;
    s->full_every = full_rebuild_every;
    if (const char* e = getenv("QCF_INC_RMS")) s->inc_rms = std::max(0.0, atof(e));
    const size_t nn = (size_t)N * N;
    CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&s->e0)); CK(cudaEventCreate(&s->e1)); CK(cudaEventCreate(&s->e2));
    CKB(cublasCreate(&s->blas));
    CKB(cublasSetStream(s->blas, s->stream));
    CKS(cusolverDnCreate(&s->sol));
    CKS(cusolverDnSetStream(s->sol, s->stream));
    for (double** p : {&s->S, &s->H, &s->X, &s->T1, &s->T2, &s->Fx, &s->C}) CK(cudaMalloc(p, nn * sizeof(double)));
    CK(cudaMalloc(&s->W, N * sizeof(double)));
    CK(cudaMalloc(&s->scal, 32 * sizeof(double)));
    CK(cudaMalloc(&s->info, sizeof(int)));
    for (int k = 0; k < s->nspin; ++k) {
        CK(cudaMalloc(&s->P[k], nn * sizeof(double))); CK(cudaMalloc(&s->G[k], nn * sizeof(double))); CK(cudaMalloc(&s->F[k], nn * sizeof(double)));
        CK(cudaMalloc(&s->Pnext[k], nn * sizeof(double)));
        s->diis[k].min_len = unrestricted ? 2 : 4;      // uhf.rs:76-78 / rhf.rs:65
        s->diis[k].max_len = unrestricted ? 8 : 6;
        s->diis[k].pool.resize(s->diis[k].max_len);
        for (auto& sm : s->diis[k].pool) { CK(cudaMalloc(&sm.err, nn * sizeof(double))); CK(cudaMalloc(&sm.fock, nn * sizeof(double))); sm.id = -1; }
        CK(cudaMemsetAsync(s->G[k], 0, nn * sizeof(double), s->stream));
    }
    CKS(cusolverDnDsyevd_bufferSize(s->sol, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, s->T2, N, s->W, &s->lwork));
    CK(cudaMalloc(&s->work, sizeof(double) * std::max(s->lwork, 1)));
    CK(cudaMemcpyAsync(s->S, S, nn * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(s->H, H, nn * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    // X = U s^-1/2 U^T (rhf.rs:124-131)
    CK(cudaMemcpyAsync(s->T2, s->S, nn * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    int rc = eigh(ctx, s, s->T2);
    if (rc) return rc;
    CK(cudaMemcpyAsync(s->T1, s->T2, nn * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    scale_columns_kernel<<<nblk(nn), 256, 0, s->stream>>>(N, s->T1, s->W);
    rc = gemm(ctx, s, CUBLAS_OP_N, CUBLAS_OP_T, N, N, N, 1.0, s->T1, s->T2, 0.0, s->X);
    if (rc) return rc;
    // Hueckel guess (rhf.rs:133-150, uhf.rs:191-208); both spins share the orbitals, occupations may differ
    hueckel_kernel<<<nblk(nn), 256, 0, s->stream>>>(N, s->S, s->H, s->Fx);
    rc = orbitals_of(ctx, s, s->Fx, s->C);
    if (rc) return rc;
    for (int k = 0; k < s->nspin; ++k) {
        rc = density_from(ctx, s, s->C, s->nocc[k], s->P[k]);
        if (rc) return rc;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s->stream));
    int info = 0;
    CK(cudaMemcpy(&info, s->info, sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0) { ctx->err = "cusolverDnDsyevd did not converge (info = " + std::to_string(info) + ")"; return QCF_ERR_CUDA; }
    s->iteration = 0;
    s->since_full = 0;
    s->last_rms = 1.0;
    return QCF_OK;
}

int qcf_scf_step(qcf_ctx* ctx, double epsilon, qcf_scf_info* out) {
    if (!ctx || !out) return QCF_ERR_ARG;
    qcf_scf* s = ctx->scf;
    if (!s) { ctx->err = "qcf_scf_step before qcf_scf_init"; return QCF_ERR_STATE; }
    CK(cudaSetDevice(ctx->dev[0].device));
    const int N = s->N;
    const size_t nn = (size_t)N * N;
    const auto t0 = std::chrono::steady_clock::now();
    CK(cudaEventRecord(s->e0, s->stream));
    // ---- G(P): full or difference-density build, enqueued on the SCF stream (rhf.rs:67-68, uhf.rs:90-91) ----
    // difference-density builds once the density moves little (their screening runs on |P - P_prev|, at tau / 8): while the
    // SCF is far from convergence a full build is cheaper; a full rebuild every `full_every` steps bounds the accumulated error
    const bool inc = s->full_every > 0;
    const bool reset = inc && (s->since_full == 0 || s->last_rms > s->inc_rms);
    int rc = qcf_internal::run_build_scf(ctx, s->nspin == 2 ? 1 : 0, s->P[0], s->nspin == 2 ? s->P[1] : nullptr, s->G[0],
                                         s->nspin == 2 ? s->G[1] : nullptr, s->stream, inc, reset);
    if (rc) return rc;
    if (inc) s->since_full = reset ? 1 % s->full_every : (s->since_full + 1) % s->full_every;
    CK(cudaEventRecord(s->e1, s->stream));
    double e_sum = 0.0, rms_sum = 0.0;
    double h[2][3];
    for (int k = 0; k < s->nspin; ++k) {
        // F = H + G ; err = F P S - S P F (rhf.rs:70-71, uhf.rs:93-94)
        add_kernel<<<nblk(nn), 256, 0, s->stream>>>(nn, s->H, s->G[k], s->F[k]);
        if ((rc = gemm(ctx, s, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, 1.0, s->F[k], s->P[k], 0.0, s->T1))) return rc;
        if ((rc = gemm(ctx, s, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, 1.0, s->T1, s->S, 0.0, s->Fx))) return rc;
        if ((rc = gemm(ctx, s, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, 1.0, s->S, s->P[k], 0.0, s->T1))) return rc;
        if ((rc = gemm(ctx, s, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, -1.0, s->T1, s->F[k], 1.0, s->Fx))) return rc;
        // DIIS extrapolation (rhf.rs:73, uhf.rs:95-97); the extrapolated matrix lands in Pnext[k] (free until below)
        if ((rc = diis_fock(ctx, s, s->diis[k], s->Fx, s->F[k], s->Pnext[k]))) return rc;
        // F' = X^T F X, eigensolve, C = X C' (rhf.rs:74-76, uhf.rs:99-104)
        if ((rc = orbitals_of(ctx, s, s->Pnext[k], s->C))) return rc;
        s->eps[k].resize(N);
        CK(cudaMemcpyAsync(s->eps[k].data(), s->W, N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        // new density (rhf.rs:78, uhf.rs:110-121); the old one stays in P[k] until both spins are done (uhf.rs:81-108)
        if ((rc = density_from(ctx, s, s->C, s->nocc[k], s->Pnext[k]))) return rc;
        // change on the diagonal (rhf.rs:87-88) and E = 1/2 tr(P_new (2H + G_old)) (rhf.rs:84-85, uhf.rs:145-151)
        diag_change_kernel<<<1, 256, 0, s->stream>>>(N, s->Pnext[k], s->P[k], s->scal + 8 + 3 * k);
        CKB(cublasSetPointerMode(s->blas, CUBLAS_POINTER_MODE_DEVICE));
        CKB(cublasDdot(s->blas, (int)nn, s->Pnext[k], 1, s->H, 1, s->scal + 9 + 3 * k));
        CKB(cublasDdot(s->blas, (int)nn, s->Pnext[k], 1, s->G[k], 1, s->scal + 10 + 3 * k));
        CKB(cublasSetPointerMode(s->blas, CUBLAS_POINTER_MODE_HOST));
        CK(cudaMemcpyAsync(h[k], s->scal + 8 + 3 * k, 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
    // density += 1.0 * change (rhf.rs:80-82): the buffers keep their addresses so that the captured build graph stays valid
    for (int k = 0; k < s->nspin; ++k)
        CK(cudaMemcpyAsync(s->P[k], s->Pnext[k], nn * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    CK(cudaEventRecord(s->e2, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    int info = 0;
    CK(cudaMemcpy(&info, s->info, sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0) { ctx->err = "cusolverDnDsyevd did not converge (info = " + std::to_string(info) + ")"; return QCF_ERR_CUDA; }
    for (int k = 0; k < s->nspin; ++k) {
        rms_sum += std::sqrt(h[k][0] / N);
        e_sum += 0.5 * (2.0 * h[k][1] + h[k][2]);
    }
    float bms = 0, sms = 0;
    CK(cudaEventElapsedTime(&bms, s->e0, s->e1));
    CK(cudaEventElapsedTime(&sms, s->e1, s->e2));
    double rms = rms_sum;
    if (s->nspin == 2) rms = rms_sum / 2.0 / 2.0;     // uhf.rs:137 and :139
    s->last_rms = rms;
    out->iteration = s->iteration;
    out->converged = rms < epsilon ? 1 : 0;
    out->electronic_energy = e_sum;
    out->density_rms = rms;
    out->build_ms = bms;
    out->linalg_ms = sms;
    out->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    ++s->iteration;
    return QCF_OK;
}

// what: 0 density, 1 Fock matrix F = H + G of the last step, 2 G of the last step (N*N each); 3 orbital energies (N)
int qcf_scf_get(qcf_ctx* ctx, int what, int spin, double* out) {
    if (!ctx || !out) return QCF_ERR_ARG;
    qcf_scf* s = ctx->scf;
    if (!s) { ctx->err = "qcf_scf_get before qcf_scf_init"; return QCF_ERR_STATE; }
    if (spin < 0 || spin >= s->nspin || what < 0 || what > 3) return QCF_ERR_ARG;
    CK(cudaSetDevice(ctx->dev[0].device));
    const size_t nn = (size_t)s->N * s->N;
    if (what == 3) {
        if ((int)s->eps[spin].size() != s->N) { ctx->err = "no SCF step has run yet"; return QCF_ERR_STATE; }
        std::copy(s->eps[spin].begin(), s->eps[spin].end(), out);
        return QCF_OK;
    }
    const double* src = what == 0 ? s->P[spin] : what == 1 ? s->F[spin] : s->G[spin];
    CK(cudaMemcpy(out, src, nn * sizeof(double), cudaMemcpyDeviceToHost));
    return QCF_OK;
}

}  // extern "C"
