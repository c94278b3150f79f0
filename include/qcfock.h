/* qcfock.h -- C ABI of the B200-native direct-SCF Fock-build engine for qchem-rs.
 *
 * This is the drop-in boundary.  The reference has no FFI today: the path is a plain
 * Rust call `molint::eri(&MolecularSystem) -> EriTensor` (core/src/hf/rhf.rs:45, uhf.rs:55) whose
 * result is repacked (rhf.rs:58-62) and contracted with the density once per SCF iteration
 * (rhf.rs:67-68 -> :152-167, uhf.rs:90-91 -> :210-227).  The engine fuses those pieces into one call
 * per iteration, G = fock_build(P), and never stores the N^4 tensor.
 *
 * Conventions
 *   - every function returns 0 on success and a negative code on failure, never throws / unwinds
 *     across the boundary; qcf_last_error() gives the text of the last failure on that context;
 *   - matrices are dense column-major N x N float64 (nalgebra `DMatrix<f64>`); they are symmetric,
 *     so row- and column-major views coincide;
 *   - the caller owns all host buffers, the library copies on entry; the library owns device memory
 *     and streams; calls on one context are not re-entrant;
 *   - there is NO CPU fallback: without a usable CUDA device qcf_create fails with QCF_ERR_CUDA.
 */
#ifndef QCFOCK_H
#define QCFOCK_H

#ifdef __cplusplus
extern "C" {
#endif

#define QCF_OK 0
#define QCF_ERR_ARG (-1)      /* bad argument (null pointer, l > 2, spherical shells, ...)   */
#define QCF_ERR_CUDA (-2)     /* CUDA runtime failure; text in qcf_last_error              */
#define QCF_ERR_STATE (-3)    /* call not valid for this context                            */

typedef struct qcf_ctx qcf_ctx; /* opaque */

/* Molecule + basis in flat form.  Replaces what the reference reaches through
 * `&MolecularSystem` (atoms: rhf.rs:36, 116-117; shells: inside molint).  Fused SP shells are already
 * split; `coefs[k]` = contraction coefficient x normalisation of x^l exp(-a r^2); the engine applies
 * the remaining per-component factor (xy vs xx).  Basis functions are shell-major in the order given,
 * Cartesian components x,y,z / xx,xy,xz,yy,yz,zz.  Positions in bohr. */
typedef struct {
    int n_atoms;
    const int* Z;              /* [n_atoms]                                    */
    const double* xyz;         /* [3*n_atoms] bohr                             */
    int n_shells;
    const int* shell_atom;     /* [n_shells] index into the atom arrays        */
    const int* shell_l;        /* [n_shells] 0, 1 or 2                         */
    const int* shell_nprim;    /* [n_shells]                                   */
    const int* shell_prim_off; /* [n_shells] first primitive in exps/coefs     */
    const double* exps;
    const double* coefs;
    int cartesian;             /* must be 1 (6d)                               */
} qcf_basis;

typedef struct {
    double screen_tau; /* skip quartets with Q_ab Q_cd D_max < tau; <= 0 selects the default 1e-12;
                          QCF_TAU_NONE (any value < -0.5) disables screening entirely            */
    int device;        /* CUDA device ordinal                                                    */
    int rank;          /* this context evaluates bra pairs i with i % world_size == rank;         */
    int world_size;    /*   the caller sums the partial results of all ranks (one allreduce)      */
    int block_threads; /* 0 = default                                                             */
} qcf_opts;
#define QCF_TAU_NONE (-1.0)

typedef struct {
    int n_basis, n_shells, n_pairs, n_groups;
    long long quartets;        /* unique shell quartets evaluated by the last build (this rank)   */
    long long quartets_total;  /* unique shell quartets before any screening (whole problem)      */
    double model_flops;        /* SURVEY.md 8d op-count model over the evaluated quartets          */
    double kernel_ms;          /* device time of the last build's kernels (CUDA events)           */
    double total_ms;           /* last build, including H2D / D2H copies                          */
    int launches;              /* kernels launched by the last build                              */
    long long prim_pairs;      /* primitive pairs of all shell pairs before primitive screening    */
    long long prim_pairs_kept; /* ... and those the kernels loop over                               */
} qcf_stats_t;

/* Build shell pairs, Schwarz bounds and the device-resident pair data.  (Replaces the one-off
 * molint::eri call, rhf.rs:45 / uhf.rs:55.) */
int qcf_create(const qcf_basis* basis, const qcf_opts* opts, qcf_ctx** out);
int qcf_nbasis(const qcf_ctx* ctx);

/* RHF: G = J[P] - K[P]/2, P carrying the factor 2 (rhf.rs:169-181).  Replaces rhf.rs:58-62 + :152-167. */
int qcf_build_rhf(qcf_ctx* ctx, const double* P, double* G);
/* UHF: Ga = J[Pa+Pb] - K[Pa], Gb = J[Pa+Pb] - K[Pb].  Replaces the two calls at uhf.rs:90-91 (:210-227). */
int qcf_build_uhf(qcf_ctx* ctx, const double* Pa, const double* Pb, double* Ga, double* Gb);
/* J_d = sum_kl P_d,kl (ij|kl),  K_d = sum_kl P_d,kl (ik|jl)  for nd densities. */
int qcf_build_jk(qcf_ctx* ctx, int nd, const double* const* P, double* const* J, double* const* K);

/* Device-resident variants for multi-process runs: dP / dG are device pointers on the context's
 * device (N*N doubles each), the call is asynchronous on `stream` (a cudaStream_t, may be 0) and the
 * result is this rank's PARTIAL, already symmetrised, matrix -- sum it over ranks with one
 * allreduce (ncclAllReduce / torch.distributed.all_reduce). */
int qcf_build_rhf_dev(qcf_ctx* ctx, const double* dP, double* dG, void* stream);
int qcf_build_uhf_dev(qcf_ctx* ctx, const double* dPa, const double* dPb, double* dGa, double* dGb, void* stream);

/* One-electron matrices S (overlap), T (kinetic), V (nuclear attraction, charges Z), N x N each.
 * Replaces molint::overlap / kinetic / nuclear (rhf.rs:41-43, uhf.rs:52-54), so that a driver linked
 * against this library no longer needs the absent molint crate. */
int qcf_one_electron(qcf_ctx* ctx, double* S, double* T, double* V);

/* Contracted two-electron integrals of one shell quartet, out[na*nb*nc*nd] row-major (ab|cd),
 * fully normalised.  Computed by the same device code as the Fock build (parity tests). */
int qcf_eri_quartet(qcf_ctx* ctx, int sa, int sb, int sc, int sd, double* out);
/* Schwarz factors Q[sa*n_shells+sb] = sqrt(max|(ab|ab)|), 0 for pairs dropped at creation. */
int qcf_schwarz(qcf_ctx* ctx, double* Q);
/* Boys function F_0..F_mmax(T) as evaluated on the device (parity tests), n values of T. */
int qcf_boys(qcf_ctx* ctx, int mmax, int n, const double* T, double* F);

/* Measured FP64 FMA peak of the context's device in TFLOP/s (dependent-chain microbenchmark). */
int qcf_fp64_peak(qcf_ctx* ctx, double* tflops);

int qcf_stats(const qcf_ctx* ctx, qcf_stats_t* out);

/* Per-launch record of the last build: class (la lb|lc ld), primitive-pair counts, list lengths,
 * unique shell quartets evaluated, model flops per primitive quartet, and -- only when the context was
 * created with the environment variable QCF_PROFILE=1, which serialises the launches -- its device time.
 * Returns the number of launches of the last build (writes at most max_rec records). */
typedef struct {
    int la, lb, kab, lc, ld, kcd, nbra, nket;
    long long quartets;
    double flops_per_prim_quartet;
    float ms;
} qcf_launch_rec;
int qcf_launch_profile(qcf_ctx* ctx, int max_rec, qcf_launch_rec* out);
const char* qcf_last_error(const qcf_ctx* ctx);
void qcf_destroy(qcf_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* QCFOCK_H */
