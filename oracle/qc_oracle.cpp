// qc_oracle.cpp -- CPU ORACLE for the Fock-build path of qchem-rs.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (qchem-rs_b200/, libqcfock.so) never links or calls it.
//
// PARITY UNPINNED BY THE REFERENCE: the reference's integral arithmetic lives in the crate
// `molint`, an un-vendored path dependency (`molint = { path = "../molint" }`, Cargo.toml:12, no
// version, no lock file) plus `boys = "0.1.0"` (core/Cargo.toml:17); neither is present under
// /root/reference and the reference ships no tests or golden vectors.  What pins this oracle
// instead (tests/test_oracle.py):
//   * the Szabo-Ostlund H2/STO-3G table (S12, T, V, (11|11)..., E_tot = -1.1167143 Eh) computed
//     from the reference's own data files,
//   * tests/golden/*.json produced by an independent closed-form (Taketa-Huzinaga-O-ohata)
//     Python implementation with scipy's 1F1 as the Boys function (tests/golden/make_golden.py),
//   * 8-fold permutational symmetry, rotation/translation invariance, tr(PS) = n_electrons,
//   * dense "reference-faithful" contraction == direct shell-quartet digestion at tau = 0.
//
// What it restates, by reference file:line:
//   orc_overlap / orc_kinetic / orc_nuclear  <- molint::overlap/kinetic/nuclear  rhf.rs:41-43
//   orc_eri_tensor                           <- molint::eri                      rhf.rs:45, uhf.rs:55
//   orc_fock_rhf_dense                       <- electron_terms build + compute_electronic_hamiltonian
//                                               rhf.rs:58-62, rhf.rs:152-167 (upper triangle, mirrored
//                                               as utils.rs:7-13)
//   orc_fock_uhf_dense                       <- uhf.rs:210-227
//   orc_fock_direct                          <- same G, direct-SCF form (the reference cannot hold
//                                               N^4 doubles for N >= ~300), OpenMP over bra pairs
// Integrals: textbook McMurchie-Davidson (Helgaker, Jorgensen, Olsen, ch. 9), Cartesian GTOs.
//
// Build: see oracle/Makefile  (g++ -O2 -fopenmp -shared -fPIC).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {
struct orc_basis {
    int n_atoms; const int* Z; const double* xyz;
    int n_shells; const int* shell_atom; const int* shell_l; const int* shell_nprim;
    const int* shell_prim_off; const double* exps; const double* coefs; int cartesian;
};
}

namespace {

constexpr int LMAX = 3;                 // oracle handles s..f
constexpr int LTOT = 4 * LMAX;          // max Boys order
const double PI = 3.14159265358979323846;

inline int ncart(int l) { return (l + 1) * (l + 2) / 2; }

double dfact(int n) { double r = 1; while (n > 1) { r *= n; n -= 2; } return r; }

// Boys function F_m(T), m = 0..mmax.  Series + downward recursion for small/medium T, erf-based F_0 +
// upward recursion for large T.  long double inside; compared with scipy hyp1f1 in tests.
void boys(int mmax, double T, double* F) {
    long double t = T;
    if (T < 35.0 + 2.0 * mmax) {
        // F_m(T) = exp(-T) * sum_k (2T)^k / ((2m+1)(2m+3)...(2m+2k+1))
        long double term = 1.0L / (2 * mmax + 1), sum = term;
        for (int k = 1; k < 400; ++k) {
            term *= 2.0L * t / (2 * mmax + 2 * k + 1);
            sum += term;
            if (term < 1e-22L * sum) break;
        }
        long double e = expl(-t);
        long double f = e * sum;
        F[mmax] = (double)f;
        for (int m = mmax; m > 0; --m) {
            f = (2.0L * t * f + e) / (2 * m - 1);
            F[m - 1] = (double)f;
        }
    } else {
        long double e = expl(-t);
        long double f = 0.5L * sqrtl((long double)PI / t) * erfl(sqrtl(t));
        F[0] = (double)f;
        for (int m = 0; m < mmax; ++m) {
            f = ((2 * m + 1) * f - e) / (2.0L * t);
            F[m + 1] = (double)f;
        }
    }
}

// Hermite expansion coefficients E_t^{ij} for one Cartesian axis (HJO eq. 9.5.6-9.5.7).
// E[i][j][t], 0<=i<=la, 0<=j<=lb, 0<=t<=i+j.  XPA = P-A, XPB = P-B, includes exp(-mu XAB^2) if k0 given.
struct ETab {
    double e[LMAX + 1][LMAX + 1][2 * LMAX + 1];
};
void hermite_E(int la, int lb, double p, double XPA, double XPB, double k0, ETab& E) {
    std::memset(&E, 0, sizeof(E));
    const double h = 0.5 / p;
    E.e[0][0][0] = k0;
    for (int i = 0; i <= la; ++i) {
        if (i > 0) {
            for (int t = 0; t <= i; ++t) {
                double v = XPA * E.e[i - 1][0][t];
                if (t > 0) v += h * E.e[i - 1][0][t - 1];
                if (t + 1 <= i - 1) v += (t + 1) * E.e[i - 1][0][t + 1];
                E.e[i][0][t] = v;
            }
        }
        for (int j = 1; j <= lb; ++j) {
            for (int t = 0; t <= i + j; ++t) {
                double v = XPB * E.e[i][j - 1][t];
                if (t > 0) v += h * E.e[i][j - 1][t - 1];
                if (t + 1 <= i + j - 1) v += (t + 1) * E.e[i][j - 1][t + 1];
                E.e[i][j][t] = v;
            }
        }
    }
}

// Hermite Coulomb integrals R_{tuv} = R^0_{tuv}(alpha, PQ), t+u+v <= L  (HJO eq. 9.9.18-9.9.20)
struct RTab {
    double r[LTOT + 1][LTOT + 1][LTOT + 1];
};
void hermite_R(int L, double alpha, double X, double Y, double Z, RTab& R) {
    static thread_local double aux[LTOT + 1][LTOT + 1][LTOT + 1][LTOT + 1];  // [n][t][u][v]
    double F[LTOT + 1];
    boys(L, alpha * (X * X + Y * Y + Z * Z), F);
    double m2a = 1.0;
    for (int n = 0; n <= L; ++n) { aux[n][0][0][0] = m2a * F[n]; m2a *= -2.0 * alpha; }
    for (int n = L - 1; n >= 0; --n) {
        int K = L - n;
        for (int t = 0; t <= K; ++t)
            for (int u = 0; t + u <= K; ++u)
                for (int v = 0; t + u + v <= K; ++v) {
                    if (t + u + v == 0) continue;
                    double val;
                    if (t > 0) {
                        val = X * aux[n + 1][t - 1][u][v];
                        if (t > 1) val += (t - 1) * aux[n + 1][t - 2][u][v];
                    } else if (u > 0) {
                        val = Y * aux[n + 1][t][u - 1][v];
                        if (u > 1) val += (u - 1) * aux[n + 1][t][u - 2][v];
                    } else {
                        val = Z * aux[n + 1][t][u][v - 1];
                        if (v > 1) val += (v - 1) * aux[n + 1][t][u][v - 2];
                    }
                    aux[n][t][u][v] = val;
                }
    }
    for (int t = 0; t <= L; ++t)
        for (int u = 0; t + u <= L; ++u)
            for (int v = 0; t + u + v <= L; ++v) R.r[t][u][v] = aux[0][t][u][v];
}

struct Comp { int x, y, z; double scale; };
std::vector<Comp> components(int l) {
    std::vector<Comp> c;
    for (int i = l; i >= 0; --i)
        for (int j = l - i; j >= 0; --j) {
            int k = l - i - j;
            c.push_back({i, j, k, std::sqrt(dfact(2 * l - 1) / (dfact(2 * i - 1) * dfact(2 * j - 1) * dfact(2 * k - 1)))});
        }
    return c;
}

// cached per l (components() allocates; the quartet routine is called ~1e7 times in the size-parity tests)
const std::vector<Comp>& components_cached(int l) {
    static const std::vector<Comp> tab[LMAX + 1] = {components(0), components(1), components(2), components(3)};
    return tab[l];
}

struct ShellView {
    int l, nprim; const double* exps; const double* coefs; double A[3];
};
ShellView shell(const orc_basis* b, int s) {
    ShellView v;
    v.l = b->shell_l[s]; v.nprim = b->shell_nprim[s];
    v.exps = b->exps + b->shell_prim_off[s]; v.coefs = b->coefs + b->shell_prim_off[s];
    const double* a = b->xyz + 3 * b->shell_atom[s];
    v.A[0] = a[0]; v.A[1] = a[1]; v.A[2] = a[2];
    return v;
}

std::vector<int> offsets(const orc_basis* b) {
    std::vector<int> off(b->n_shells + 1, 0);
    for (int s = 0; s < b->n_shells; ++s) off[s + 1] = off[s] + ncart(b->shell_l[s]);
    return off;
}

// Contracted ERI block (ab|cd), out[na][nb][nc][nd] row-major, chemists' notation.
void eri_quartet(const ShellView& A, const ShellView& B, const ShellView& C, const ShellView& D, double* out) {
    const auto &ca = components_cached(A.l), &cb = components_cached(B.l), &cc = components_cached(C.l), &cd = components_cached(D.l);
    const int na = ca.size(), nb = cb.size(), nc = cc.size(), nd = cd.size();
    std::fill(out, out + (size_t)na * nb * nc * nd, 0.0);
    const int L = A.l + B.l + C.l + D.l;
    double AB2 = 0, CD2 = 0;
    for (int k = 0; k < 3; ++k) { AB2 += (A.A[k] - B.A[k]) * (A.A[k] - B.A[k]); CD2 += (C.A[k] - D.A[k]) * (C.A[k] - D.A[k]); }
    ETab Eab[3], Ecd[3];
    RTab R;
    for (int ia = 0; ia < A.nprim; ++ia)
    for (int ib = 0; ib < B.nprim; ++ib) {
        const double a = A.exps[ia], b = B.exps[ib], p = a + b;
        double P[3];
        for (int k = 0; k < 3; ++k) P[k] = (a * A.A[k] + b * B.A[k]) / p;
        const double Kab = std::exp(-a * b / p * AB2) * A.coefs[ia] * B.coefs[ib];
        for (int k = 0; k < 3; ++k) hermite_E(A.l, B.l, p, P[k] - A.A[k], P[k] - B.A[k], 1.0, Eab[k]);
        for (int ic = 0; ic < C.nprim; ++ic)
        for (int id = 0; id < D.nprim; ++id) {
            const double c = C.exps[ic], d = D.exps[id], q = c + d;
            double Q[3];
            for (int k = 0; k < 3; ++k) Q[k] = (c * C.A[k] + d * D.A[k]) / q;
            const double Kcd = std::exp(-c * d / q * CD2) * C.coefs[ic] * D.coefs[id];
            for (int k = 0; k < 3; ++k) hermite_E(C.l, D.l, q, Q[k] - C.A[k], Q[k] - D.A[k], 1.0, Ecd[k]);
            const double alpha = p * q / (p + q);
            hermite_R(L, alpha, P[0] - Q[0], P[1] - Q[1], P[2] - Q[2], R);
            const double pref = 2.0 * std::pow(PI, 2.5) / (p * q * std::sqrt(p + q)) * Kab * Kcd;
            size_t idx = 0;
            for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j)
            for (int k = 0; k < nc; ++k) for (int l = 0; l < nd; ++l, ++idx) {
                double sum = 0;
                for (int t = 0; t <= ca[i].x + cb[j].x; ++t)
                for (int u = 0; u <= ca[i].y + cb[j].y; ++u)
                for (int v = 0; v <= ca[i].z + cb[j].z; ++v) {
                    const double eab = Eab[0].e[ca[i].x][cb[j].x][t] * Eab[1].e[ca[i].y][cb[j].y][u] * Eab[2].e[ca[i].z][cb[j].z][v];
                    double inner = 0;
                    for (int tt = 0; tt <= cc[k].x + cd[l].x; ++tt)
                    for (int uu = 0; uu <= cc[k].y + cd[l].y; ++uu)
                    for (int vv = 0; vv <= cc[k].z + cd[l].z; ++vv) {
                        const double ecd = Ecd[0].e[cc[k].x][cd[l].x][tt] * Ecd[1].e[cc[k].y][cd[l].y][uu] * Ecd[2].e[cc[k].z][cd[l].z][vv];
                        const double sgn = ((tt + uu + vv) & 1) ? -1.0 : 1.0;
                        inner += sgn * ecd * R.r[t + tt][u + uu][v + vv];
                    }
                    sum += eab * inner;
                }
                out[idx] += pref * sum * ca[i].scale * cb[j].scale * cc[k].scale * cd[l].scale;
            }
        }
    }
}

// one-electron blocks ---------------------------------------------------------------------------
// kind 0 overlap, 1 kinetic, 2 nuclear attraction
void one_electron_block(const orc_basis* bs, const ShellView& A, const ShellView& B, int kind, double* out) {
    const auto ca = components(A.l), cb = components(B.l);
    const int na = ca.size(), nb = cb.size();
    std::fill(out, out + na * nb, 0.0);
    double AB2 = 0;
    for (int k = 0; k < 3; ++k) AB2 += (A.A[k] - B.A[k]) * (A.A[k] - B.A[k]);
    for (int ia = 0; ia < A.nprim; ++ia)
    for (int ib = 0; ib < B.nprim; ++ib) {
        const double a = A.exps[ia], b = B.exps[ib], p = a + b;
        double P[3];
        for (int k = 0; k < 3; ++k) P[k] = (a * A.A[k] + b * B.A[k]) / p;
        const double Kab = std::exp(-a * b / p * AB2) * A.coefs[ia] * B.coefs[ib];
        if (kind == 2) {
            ETab E[3];
            for (int k = 0; k < 3; ++k) hermite_E(A.l, B.l, p, P[k] - A.A[k], P[k] - B.A[k], 1.0, E[k]);
            RTab R;
            for (int at = 0; at < bs->n_atoms; ++at) {
                const double* C = bs->xyz + 3 * at;
                hermite_R(A.l + B.l, p, P[0] - C[0], P[1] - C[1], P[2] - C[2], R);
                const double pref = -bs->Z[at] * 2.0 * PI / p * Kab;
                for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j) {
                    double sum = 0;
                    for (int t = 0; t <= ca[i].x + cb[j].x; ++t)
                    for (int u = 0; u <= ca[i].y + cb[j].y; ++u)
                    for (int v = 0; v <= ca[i].z + cb[j].z; ++v)
                        sum += E[0].e[ca[i].x][cb[j].x][t] * E[1].e[ca[i].y][cb[j].y][u] * E[2].e[ca[i].z][cb[j].z][v] * R.r[t][u][v];
                    out[i * nb + j] += pref * sum * ca[i].scale * cb[j].scale;
                }
            }
        } else {
            // overlap factors S_ij per axis with j up to lb+2 for the kinetic operator (HJO 9.3.x)
            static thread_local double e[3][LMAX + 1][LMAX + 3];
            for (int k = 0; k < 3; ++k) {
                // E_0^{ij} by the same two-term recursion on t=0 needs the full table; reuse hermite_E on la, lb+2
                // with a local larger table:
                double tab[LMAX + 1][LMAX + 3][2 * LMAX + 3];
                std::memset(tab, 0, sizeof(tab));
                const double h = 0.5 / p, XPA = P[k] - A.A[k], XPB = P[k] - B.A[k];
                tab[0][0][0] = 1.0;
                for (int i = 0; i <= A.l; ++i) {
                    if (i > 0) for (int t = 0; t <= i; ++t) {
                        double v = XPA * tab[i - 1][0][t];
                        if (t > 0) v += h * tab[i - 1][0][t - 1];
                        if (t + 1 <= i - 1) v += (t + 1) * tab[i - 1][0][t + 1];
                        tab[i][0][t] = v;
                    }
                    for (int j = 1; j <= B.l + 2; ++j) for (int t = 0; t <= i + j; ++t) {
                        double v = XPB * tab[i][j - 1][t];
                        if (t > 0) v += h * tab[i][j - 1][t - 1];
                        if (t + 1 <= i + j - 1) v += (t + 1) * tab[i][j - 1][t + 1];
                        tab[i][j][t] = v;
                    }
                }
                for (int i = 0; i <= A.l; ++i) for (int j = 0; j <= B.l + 2; ++j) e[k][i][j] = tab[i][j][0];
            }
            const double s0 = std::pow(PI / p, 1.5) * Kab;
            for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j) {
                const int ax[3] = {ca[i].x, ca[i].y, ca[i].z}, bx[3] = {cb[j].x, cb[j].y, cb[j].z};
                double val;
                if (kind == 0) {
                    val = e[0][ax[0]][bx[0]] * e[1][ax[1]][bx[1]] * e[2][ax[2]][bx[2]];
                } else {
                    // T = -1/2 <a| d2/dx2 + d2/dy2 + d2/dz2 |b>, acting on |b>:
                    // d2/dx2 x^j e^{-b x^2} = j(j-1) x^{j-2} - 2b(2j+1) x^j + 4b^2 x^{j+2}
                    val = 0;
                    for (int k = 0; k < 3; ++k) {
                        const int j0 = bx[k];
                        double d2 = 4.0 * b * b * e[k][ax[k]][j0 + 2] - 2.0 * b * (2 * j0 + 1) * e[k][ax[k]][j0];
                        if (j0 >= 2) d2 += j0 * (j0 - 1) * e[k][ax[k]][j0 - 2];
                        double term = d2;
                        for (int m = 0; m < 3; ++m) if (m != k) term *= e[m][ax[m]][bx[m]];
                        val += -0.5 * term;
                    }
                }
                out[i * nb + j] += s0 * val * ca[i].scale * cb[j].scale;
            }
        }
    }
}

void one_electron(const orc_basis* b, int kind, double* M) {
    const auto off = offsets(b);
    const int N = off.back();
    #pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < b->n_shells; ++s) {
        std::vector<double> blk(ncart(LMAX) * ncart(LMAX));
        for (int r = 0; r <= s; ++r) {
            ShellView A = shell(b, s), B = shell(b, r);
            one_electron_block(b, A, B, kind, blk.data());
            const int na = ncart(A.l), nb = ncart(B.l);
            for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j) {
                M[(size_t)(off[s] + i) * N + off[r] + j] = blk[i * nb + j];
                M[(size_t)(off[r] + j) * N + off[s] + i] = blk[i * nb + j];
            }
        }
    }
}

}  // namespace

extern "C" {

int orc_nbasis(const orc_basis* b) { return offsets(b).back(); }

void orc_boys(int mmax, double T, double* F) { boys(mmax, T, F); }

void orc_overlap(const orc_basis* b, double* S) { one_electron(b, 0, S); }
void orc_kinetic(const orc_basis* b, double* T) { one_electron(b, 1, T); }
void orc_nuclear(const orc_basis* b, double* V) { one_electron(b, 2, V); }

// rhf.rs:110-122
double orc_nuclear_repulsion(const orc_basis* b) {
    double e = 0;
    for (int i = 0; i < b->n_atoms; ++i)
        for (int j = i + 1; j < b->n_atoms; ++j) {
            double r2 = 0;
            for (int k = 0; k < 3; ++k) { double d = b->xyz[3 * j + k] - b->xyz[3 * i + k]; r2 += d * d; }
            e += (double)(b->Z[i] * b->Z[j]) / std::sqrt(r2);
        }
    return e;
}

// one contracted shell quartet block, out[na*nb*nc*nd]
void orc_eri_shell_quartet(const orc_basis* b, int sa, int sb, int sc, int sd, double* out) {
    eri_quartet(shell(b, sa), shell(b, sb), shell(b, sc), shell(b, sd), out);
}

// Full N^4 tensor, eri[((i*N+j)*N+k)*N+l] = (ij|kl)   (molint::eri, rhf.rs:45)
void orc_eri_tensor(const orc_basis* b, double* eri) {
    const auto off = offsets(b);
    const size_t N = off.back();
    const int ns = b->n_shells;
    #pragma omp parallel for schedule(dynamic) collapse(2)
    for (int sa = 0; sa < ns; ++sa)
    for (int sb = 0; sb < ns; ++sb) {
        if (sb > sa) continue;
        std::vector<double> blk((size_t)ncart(LMAX) * ncart(LMAX) * ncart(LMAX) * ncart(LMAX));
        for (int sc = 0; sc <= sa; ++sc)
        for (int sd = 0; sd <= sc; ++sd) {
            if (sc == sa && sd > sb) continue;
            ShellView A = shell(b, sa), B = shell(b, sb), C = shell(b, sc), D = shell(b, sd);
            eri_quartet(A, B, C, D, blk.data());
            const int na = ncart(A.l), nb = ncart(B.l), nc = ncart(C.l), nd = ncart(D.l);
            size_t idx = 0;
            for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j)
            for (int k = 0; k < nc; ++k) for (int l = 0; l < nd; ++l, ++idx) {
                const size_t I = off[sa] + i, J = off[sb] + j, K = off[sc] + k, Lx = off[sd] + l;
                const double v = blk[idx];
                eri[((I * N + J) * N + K) * N + Lx] = v; eri[((J * N + I) * N + K) * N + Lx] = v;
                eri[((I * N + J) * N + Lx) * N + K] = v; eri[((J * N + I) * N + Lx) * N + K] = v;
                eri[((K * N + Lx) * N + I) * N + J] = v; eri[((Lx * N + K) * N + I) * N + J] = v;
                eri[((K * N + Lx) * N + J) * N + I] = v; eri[((Lx * N + K) * N + J) * N + I] = v;
            }
        }
    }
}

// Reference-faithful RHF two-electron matrix: ET[ijkl] = (ij|kl) - 1/2 (ik|jl)  (rhf.rs:58-62), then
// G_ij = sum_kl P_kl ET[ijkl] for i <= j, mirrored (rhf.rs:152-167 + utils.rs:7-13).  Single thread,
// like the reference.  `et` may be NULL (allocated internally) or a caller-kept N^4 scratch that is
// filled when *et_ready == 0.
void orc_fock_rhf_dense(int n, const double* P, const double* eri, double* et, int* et_ready, double* G) {
    const size_t N = n;
    std::vector<double> local;
    if (!et) { local.resize(N * N * N * N); et = local.data(); }
    if (!et_ready || !*et_ready) {
        for (size_t i = 0; i < N; ++i) for (size_t j = 0; j < N; ++j)
        for (size_t k = 0; k < N; ++k) for (size_t l = 0; l < N; ++l)
            et[((i * N + j) * N + k) * N + l] = eri[((i * N + j) * N + k) * N + l] - 0.5 * eri[((i * N + k) * N + j) * N + l];
        if (et_ready) *et_ready = 1;
    }
    for (size_t i = 0; i < N; ++i)
        for (size_t j = i; j < N; ++j) {
            double sum = 0;
            const double* row = et + (i * N + j) * N * N;
            for (size_t k = 0; k < N; ++k) for (size_t l = 0; l < N; ++l) sum += P[k + l * N] * row[k * N + l];
            G[i + j * N] = sum; G[j + i * N] = sum;
        }
}

// uhf.rs:210-227: G_ij = sum_kl P1_kl (ij|kl) + P2_kl (ij|kl) - P1_kl (ik|jl), i <= j mirrored
void orc_fock_uhf_dense(int n, const double* P1, const double* P2, const double* eri, double* G) {
    const size_t N = n;
    for (size_t i = 0; i < N; ++i)
        for (size_t j = i; j < N; ++j) {
            double sum = 0;
            for (size_t k = 0; k < N; ++k) for (size_t l = 0; l < N; ++l)
                sum += P1[k + l * N] * eri[((i * N + j) * N + k) * N + l] + P2[k + l * N] * eri[((i * N + j) * N + k) * N + l]
                     - P1[k + l * N] * eri[((i * N + k) * N + j) * N + l];
            G[i + j * N] = sum; G[j + i * N] = sum;
        }
}

// Schwarz factors Q[sa*ns+sb] = sqrt(max |(ab|ab)|)
void orc_schwarz(const orc_basis* b, double* Q) {
    const int ns = b->n_shells;
    #pragma omp parallel for schedule(dynamic)
    for (int sa = 0; sa < ns; ++sa) {
        std::vector<double> blk((size_t)ncart(LMAX) * ncart(LMAX) * ncart(LMAX) * ncart(LMAX));
        for (int sb = 0; sb <= sa; ++sb) {
            ShellView A = shell(b, sa), B = shell(b, sb);
            eri_quartet(A, B, A, B, blk.data());
            const int na = ncart(A.l), nb = ncart(B.l);
            double m = 0;
            for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j)
                m = std::max(m, std::fabs(blk[((size_t)(i * nb + j) * na + i) * nb + j]));
            Q[sa * ns + sb] = Q[sb * ns + sa] = std::sqrt(m);
        }
    }
}

// Direct-SCF J/K.  nd densities (column-major N x N, symmetric).  J[d], K[d] are full symmetric N x N:
//   J_ij = sum_kl P_kl (ij|kl),   K_ij = sum_kl P_kl (ik|jl).
// Unique shell quartets (sa>=sb, sc>=sd, ab>=cd), Schwarz screening Q_ab Q_cd Dmax < tau skipped
// (tau = 0 -> nothing skipped).  OpenMP over bra pairs; returns the number of quartets evaluated.
// bra_stride > 1 restricts the work to bra pairs ip with ip % bra_stride == bra_offset (a bounded,
// representative sample of the same workload for bench timing; J/K are then partial).
long long orc_jk_direct(const orc_basis* b, double tau, int nd, const double* const* P, double* const* J, double* const* K,
                        const double* Q, int bra_stride, int bra_offset) {
    const auto off = offsets(b);
    const size_t N = off.back();
    const int ns = b->n_shells;
    std::vector<double> dmax((size_t)ns * ns, 0.0);
    double dglob = 0;
    for (int sa = 0; sa < ns; ++sa) for (int sb = 0; sb < ns; ++sb) {
        double m = 0;
        for (int d = 0; d < nd; ++d)
            for (int i = off[sa]; i < off[sa + 1]; ++i) for (int j = off[sb]; j < off[sb + 1]; ++j)
                m = std::max(m, std::fabs(P[d][i + j * N]));
        dmax[(size_t)sa * ns + sb] = m; dglob = std::max(dglob, m);
    }
    std::vector<std::pair<int, int>> pairs;
    for (int sa = 0; sa < ns; ++sa) for (int sb = 0; sb <= sa; ++sb) pairs.push_back({sa, sb});
    const long long npair = pairs.size();
    if (bra_stride < 1) bra_stride = 1;
    for (int d = 0; d < nd; ++d) { std::fill(J[d], J[d] + N * N, 0.0); std::fill(K[d], K[d] + N * N, 0.0); }
    long long nq = 0;
    #pragma omp parallel reduction(+ : nq)
    {
        std::vector<std::vector<double>> Jl(nd, std::vector<double>(N * N, 0.0)), Kl(nd, std::vector<double>(N * N, 0.0));
        std::vector<double> blk((size_t)ncart(LMAX) * ncart(LMAX) * ncart(LMAX) * ncart(LMAX));
        #pragma omp for schedule(dynamic, 1)
        for (long long ip = 0; ip < npair; ++ip) {
            if (ip % bra_stride != bra_offset) continue;
            const int sa = pairs[ip].first, sb = pairs[ip].second;
            const double qab = Q ? Q[sa * ns + sb] : 1.0;
            for (int sc = 0; sc <= sa; ++sc)
            for (int sd = 0; sd <= sc; ++sd) {
                if (sc == sa && sd > sb) continue;
                if (Q && tau > 0) {
                    const double qq = qab * Q[sc * ns + sd];
                    if (qq * dglob < tau) continue;
                    double dm = std::max({dmax[(size_t)sa * ns + sb], dmax[(size_t)sc * ns + sd],
                                          0.5 * dmax[(size_t)sa * ns + sc], 0.5 * dmax[(size_t)sa * ns + sd],
                                          0.5 * dmax[(size_t)sb * ns + sc], 0.5 * dmax[(size_t)sb * ns + sd]});
                    if (qq * dm < tau) continue;
                }
                ++nq;
                ShellView A = shell(b, sa), B = shell(b, sb), C = shell(b, sc), D = shell(b, sd);
                eri_quartet(A, B, C, D, blk.data());
                double deg = 1.0;
                if (sa == sb) deg *= 0.5;
                if (sc == sd) deg *= 0.5;
                if (sa == sc && sb == sd) deg *= 0.5;
                const int na = ncart(A.l), nb = ncart(B.l), nc = ncart(C.l), nD = ncart(D.l);
                size_t idx = 0;
                for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j)
                for (int k = 0; k < nc; ++k) for (int l = 0; l < nD; ++l, ++idx) {
                    const size_t a = off[sa] + i, bb = off[sb] + j, c = off[sc] + k, dd = off[sd] + l;
                    const double v = blk[idx] * deg;
                    for (int d = 0; d < nd; ++d) {
                        const double* Pd = P[d];
                        double* Jd = Jl[d].data(); double* Kd = Kl[d].data();
                        // all 8 permutations of (ab|cd), each adding to the non-symmetrised accumulators
                        Jd[a + bb * N] += 2.0 * Pd[c + dd * N] * v;  Jd[bb + a * N] += 2.0 * Pd[c + dd * N] * v;
                        Jd[c + dd * N] += 2.0 * Pd[a + bb * N] * v;  Jd[dd + c * N] += 2.0 * Pd[a + bb * N] * v;
                        Kd[a + c * N] += Pd[bb + dd * N] * v;  Kd[c + a * N] += Pd[dd + bb * N] * v;
                        Kd[a + dd * N] += Pd[bb + c * N] * v;  Kd[dd + a * N] += Pd[c + bb * N] * v;
                        Kd[bb + c * N] += Pd[a + dd * N] * v;  Kd[c + bb * N] += Pd[dd + a * N] * v;
                        Kd[bb + dd * N] += Pd[a + c * N] * v;  Kd[dd + bb * N] += Pd[c + a * N] * v;
                    }
                }
            }
        }
        #pragma omp critical
        for (int d = 0; d < nd; ++d)
            for (size_t i = 0; i < N * N; ++i) { J[d][i] += Jl[d][i]; K[d][i] += Kl[d][i]; }
    }
    return nq;
}

// Exact, UNSCREENED blocks of J and K for chosen shell pairs (sa[i], sb[i]):
//   J_ab = sum_{cd} P_cd (ab|cd),   K_ab = sum_{cd} P_cd (ac|bd)      (all c, d -- nothing is skipped)
// i.e. the reference's dense contraction (rhf.rs:58-62 + 152-167, uhf.rs:210-227) restricted to the rows/columns of
// one shell block.  This is how parity is checked at sizes where neither the N^4 tensor (554 GB at N = 513) nor a
// full unscreened direct build (1e10 quartets at N = 1007) is affordable: ~n_shell^2 quartets per block.
// Blocks are written back to back, row-major [na][nb], in the order of the pair list.
void orc_jk_blocks_exact(const orc_basis* b, const double* P, int npairs, const int* sa_list, const int* sb_list,
                         double* J, double* K) {
    const auto off = offsets(b);
    const size_t N = off.back();
    const int ns = b->n_shells;
    std::vector<size_t> boff(npairs + 1, 0);
    for (int i = 0; i < npairs; ++i)
        boff[i + 1] = boff[i] + (size_t)ncart(b->shell_l[sa_list[i]]) * ncart(b->shell_l[sb_list[i]]);
    std::fill(J, J + boff[npairs], 0.0);
    std::fill(K, K + boff[npairs], 0.0);
    const long long work = (long long)npairs * ns;
    #pragma omp parallel
    {
        std::vector<double> blk((size_t)ncart(LMAX) * ncart(LMAX) * ncart(LMAX) * ncart(LMAX));
        std::vector<double> jl(ncart(LMAX) * ncart(LMAX)), kl(ncart(LMAX) * ncart(LMAX));
        #pragma omp for schedule(dynamic, 4)
        for (long long w = 0; w < work; ++w) {
            const int ip = (int)(w / ns), sc = (int)(w % ns);
            const int sa = sa_list[ip], sb = sb_list[ip];
            ShellView A = shell(b, sa), B = shell(b, sb), C = shell(b, sc);
            const int na = ncart(A.l), nb = ncart(B.l), nc = ncart(C.l);
            std::fill(jl.begin(), jl.end(), 0.0);
            std::fill(kl.begin(), kl.end(), 0.0);
            for (int sd = 0; sd < ns; ++sd) {
                ShellView D = shell(b, sd);
                const int nd = ncart(D.l);
                if (sd <= sc) {   // Coulomb: (ab|cd), unordered shell pair {c,d} once, weight 2 off the diagonal
                    eri_quartet(A, B, C, D, blk.data());
                    const double wgt = sd == sc ? 1.0 : 2.0;
                    size_t idx = 0;
                    for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j)
                    for (int k = 0; k < nc; ++k) for (int l = 0; l < nd; ++l, ++idx)
                        jl[i * nb + j] += wgt * P[(off[sc] + k) + (off[sd] + l) * N] * blk[idx];
                }
                // exchange: (ac|bd)
                eri_quartet(A, C, B, D, blk.data());
                size_t idx = 0;
                for (int i = 0; i < na; ++i) for (int k = 0; k < nc; ++k)
                for (int j = 0; j < nb; ++j) for (int l = 0; l < nd; ++l, ++idx)
                    kl[i * nb + j] += P[(off[sc] + k) + (off[sd] + l) * N] * blk[idx];
            }
            #pragma omp critical
            for (int i = 0; i < na * nb; ++i) { J[boff[ip] + i] += jl[i]; K[boff[ip] + i] += kl[i]; }
        }
    }
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

int orc_num_procs() {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
