/*
 * md_oracle.c -- CPU ORACLE for the MolecularDynamics.jl per-step hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (moleculardynamics.jl_b200/) never links, imports or calls anything in oracle/.
 *
 * PARITY UNPINNED BY THE REFERENCE'S OWN TESTS: the reference ships no tests, fixtures or
 * golden vectors and cannot be executed here (no Julia, F4), and its pair
 * enumeration lives in the un-vendored, un-pinned CellListMap.jl (Project.toml:7).  This file
 * is therefore a literal plain-C restatement of the reference formulas, pinned instead by
 * (1) analytic known-answer values, (2) algebraic invariants and (3) an O(N^2) brute-force
 * minimum-image enumeration (tests/test_oracle_*.py).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Arrays are AoS [n][dim] doubles -- the memory image of Julia's Vector{MVector{dim,Float64}}
 * payloads / reinterpret(Float64, ...) (src/initialization.jl:44,93,97).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off; the timing variant adds -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* potential tags == include/mdb200.h MDB_POT_* */
enum { ORC_POT_PHS = 0, ORC_POT_LJ = 1, ORC_POT_XPLOR = 2, ORC_POT_POLY = 3, ORC_POT_SOFT = 4 };

/* ------------------------------------------------------------------------------------------
 * Potentials.  Each returns (u, f) like the reference's `evaluate`; the int return is 1 when the
 * potential's own range test passed (the pair "interacts"), 0 when it returned (0,0) by range.
 * ---------------------------------------------------------------------------------------- */

/* src/potentials.jl:2-3 */
static const double b_param = 1.0204081632653061;
static const double a_param = 134.5526623421209;

/* src/potentials.jl:11-29: evaluate(::PseudoHS) -> pseudohs(r, sigma; lambda=50.0).
 * lambda is a run-time variable there, so FastPow leaves `^` alone and Julia evaluates
 * Float64^Float64 (<1 ulp); libm pow is the stand-in. Note the absolute (unscaled) cut `rij < b_param`. */
static int pot_phs(double rij, double s1, double s2, double *u, double *f)
{
    const double lambda = 50.0;
    double sigma = (s1 + s2) / 2.0;
    double uij = 0.0, fij = 0.0;
    int in = 0;
    if (rij < b_param) {
        uij = a_param * (pow(sigma / rij, lambda) - pow(sigma / rij, lambda - 1.0));
        uij += 1.0;
        fij = lambda * pow(sigma / rij, lambda + 1.0);
        fij -= (lambda - 1.0) * pow(sigma / rij, lambda);
        fij *= a_param;
        in = 1;
    }
    *u = uij; *f = fij;
    return in;
}

/* src/potentials.jl:160-164 -> lj_unshifted :66-77 (shift/force_shift/V_cut/F_cut are dead).
 * p[0]=epsilon, p[1]=r_cut.  @fastpow: sr^2 -> sr*sr, sr2^3 -> (sr2*sr2)*sr2, sr6^2 -> sr6*sr6. */
static int pot_lj(const double *p, double r, double s1, double s2, double *u, double *f)
{
    double epsilon = p[0], r_cut = p[1];
    double sigma = (s1 + s2) / 2.0;
    if (r >= r_cut) { *u = 0.0; *f = 0.0; return 0; }
    double sr = sigma / r;
    double sr2 = sr * sr;
    double sr6 = (sr2 * sr2) * sr2;
    double sr12 = sr6 * sr6;
    *u = 4.0 * epsilon * (sr12 - sr6);
    *f = 24.0 * epsilon * (2.0 * sr12 - sr6) / r;
    return 1;
}

/* src/potentials.jl:190-209 xplor_switch, bug-for-bug (the first two dnum1 terms cancel). */
static void xplor_switch(double r, double r_on, double r_cut, double *S, double *dS)
{
    if (r < r_on) { *S = 1.0; *dS = 0.0; return; }
    if (r < r_cut) {
        double rc2 = r_cut * r_cut, r2 = r * r, ron2 = r_on * r_on;
        double t = rc2 - ron2;
        double denom = (t * t) * t;
        double a = rc2 - r2;
        double b = rc2 + 2.0 * r2 - 3.0 * ron2;
        double num1 = (a * a) * b;
        *S = num1 / denom;
        double dnum1 = -4.0 * r * a * b + 2.0 * a * 2.0 * r * b + (a * a) * 4.0 * r;
        *dS = dnum1 / denom;
        return;
    }
    *S = 0.0; *dS = 0.0;
}

/* src/potentials.jl:244-249 -> lj_xplor :217-236.  p[0]=epsilon, p[1]=r_on, p[2]=r_cut. */
static int pot_xplor(const double *p, double r, double s1, double s2, double *u, double *f)
{
    double eps = p[0], r_on = p[1], r_cut = p[2];
    double sigma = (s1 + s2) / 2.0;
    if (r >= r_cut) { *u = 0.0; *f = 0.0; return 0; }
    double sr = sigma / r;
    double sr2 = sr * sr;
    double sr6 = (sr2 * sr2) * sr2;
    double sr12 = sr6 * sr6;
    double V = 4.0 * eps * (sr12 - sr6);
    double F = 24.0 * eps * (2.0 * sr12 - sr6) / r;
    double S, dS;
    xplor_switch(r, r_on, r_cut, &S, &dS);
    *f = S * F + V * dS;
    *u = V * S;
    return 1;
}

/* integer powers by the multiply chains @fastpow would emit (exact chain unknown; ulp-level) */
static inline double ipow12(double x) { double x2 = x * x, x3 = x2 * x, x6 = x3 * x3; return x6 * x6; }
static inline double ipow13(double x) { return ipow12(x) * x; }
static inline double ipow14(double x) { double x2 = x * x; return ipow12(x) * x2; }
static inline double ipow16(double x) { double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4; return x8 * x8; }

/* README.md:89-145: user plugin `Polydisperse` / poly_potential.  p[0]=rcut (1.25), p[1]=non_additivity (0.2). */
/* overlap-removal penalty of the packer (stands in for Packmol.pack_monoatomic!, src/initialization.jl:20-30; the
 * package is an un-vendored dependency): u = k/2 (1 - r/tol)^2 for r < tol, f = -du/dr; p = {k, tol} */
static int pot_soft(const double *p, double r, double *u, double *f)
{
    double k = p[0], tol = p[1];
    if (!(r < tol)) { *u = 0.0; *f = 0.0; return 0; }
    double t = 1.0 - r / tol;
    *u = 0.5 * k * (t * t);
    *f = k * t / tol;
    return 1;
}

static int pot_poly(const double *p, double r, double s1, double s2, double *u, double *f)
{
    double r_cut = p[0], non_add = p[1];
    double sigma = 0.5 * (s1 + s2);
    sigma *= (1.0 - non_add * fabs(s1 - s2));
    double uij = 0.0, fij = 0.0;
    int in = 0;
    if (r < r_cut * sigma) {
        double term_1 = ipow12(sigma / r);
        double c0 = -28.0 / ipow12(r_cut);
        double c2 = 48.0 / ipow14(r_cut);
        double c4 = -21.0 / ipow16(r_cut);
        double q = r / sigma;
        double term_2 = c2 * (q * q);
        double term_3 = c4 * ((q * q) * (q * q));
        uij = term_1 + c0 + term_2 + term_3;
        fij = 12.0 * ipow12(sigma) / ipow13(r) - 2.0 * c2 * r / (sigma * sigma)
              - 4.0 * c4 * ((r * r) * r) / ((sigma * sigma) * (sigma * sigma));
        in = 1;
    }
    *u = uij; *f = fij;
    return in;
}

/* dispatch = the `evaluate(pot, d, sigma_i, sigma_j)` call at src/pairwise.jl:31 */
ORC_API int orc_evaluate(int tag, const double *p, double r, double s1, double s2, double *u, double *f)
{
    switch (tag) {
    case ORC_POT_PHS: return pot_phs(r, s1, s2, u, f);
    case ORC_POT_LJ: return pot_lj(p, r, s1, s2, u, f);
    case ORC_POT_XPLOR: return pot_xplor(p, r, s1, s2, u, f);
    case ORC_POT_POLY: return pot_poly(p, r, s1, s2, u, f);
    case ORC_POT_SOFT: return pot_soft(p, r, u, f);
    }
    *u = NAN; *f = NAN;
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * Pair enumeration (restates CellListMap.map_pairwise! as used at src/simulation.jl:100-104):
 * every unordered pair with minimum-image d2 <= cutoff^2 exactly once; coordinates are first
 * wrapped into the cell; the pair vector r = x - y is the minimum-image separation.
 * Distance formula (shared bit-for-bit with the CUDA kernels):
 *   xs = wrapped coordinate in [0, L];  dx = xs_i - xs_j;  dx = dx - k*L, k = nearbyint(dx/L);
 *   d2 = fma(dz,dz, fma(dy,dy, dx*dx)).
 * ---------------------------------------------------------------------------------------- */
static inline double wrap_coord(double x, double L, double invL)
{
    double xs = x - L * floor(x * invL);
    if (xs < 0.0) xs += L;
    if (xs >= L) xs -= L;
    return xs;
}

typedef struct {
    long double e, w;
    int64_t n_cut, n_int;
} pair_acc;

/* src/pairwise.jl:26-39 energy_and_forces!, operation order kept: (f*r_k)/d, dot(s, r). */
static inline void pair_update(int dim, const double *r, double d2, int64_t i, int64_t j, const double *diam,
                               int tag, const double *p, long double *F, pair_acc *acc, int32_t *nbr)
{
    double d = sqrt(d2);
    double u, f;
    int in = orc_evaluate(tag, p, d, diam[i], diam[j], &u, &f);
    double s[3], dot = 0.0;
    for (int k = 0; k < dim; k++) {
        s[k] = (f * r[k]) / d;
        dot = (k == 0) ? s[k] * r[k] : dot + s[k] * r[k];
    }
    acc->w += dot;
    acc->e += u;
    for (int k = 0; k < dim; k++) {
        F[i * dim + k] += s[k];
        F[j * dim + k] -= s[k];
    }
    acc->n_cut += 1;
    acc->n_int += (in == 1);
    if (nbr) { nbr[i] += 1; nbr[j] += 1; }
}

static inline int min_image(int dim, const double *xi, const double *xj, const double *box, const double *inv,
                            double cutoff2, double *r, double *d2out)
{
    double d2 = 0.0;
    for (int k = 0; k < dim; k++) {
        double dx = xi[k] - xj[k];
        double kk = nearbyint(dx * inv[k]);
        dx = dx - kk * box[k];
        r[k] = dx;
        d2 = (k == 0) ? dx * dx : fma(dx, dx, d2);
    }
    *d2out = d2;
    return d2 <= cutoff2;
}

static void finish(int dim, int64_t n, const long double *F, const pair_acc *acc, double *Fout, double *E, double *W,
                   int64_t *n_cut, int64_t *n_int)
{
    if (Fout) for (int64_t q = 0; q < n * dim; q++) Fout[q] = (double)F[q];
    if (E) *E = (double)acc->e;
    if (W) *W = (double)acc->w;
    if (n_cut) *n_cut = acc->n_cut;
    if (n_int) *n_int = acc->n_int;
}

/* O(N^2) reference enumeration: pins pair counts and forces for the cell-list oracle below. */
ORC_API int orc_forces_brute(int dim, int64_t n, const double *x, const double *diam, const double *box, double cutoff,
                             int tag, const double *p, double *Fout, double *E, double *W, int64_t *n_cut,
                             int64_t *n_int, int32_t *nbr)
{
    double inv[3];
    for (int k = 0; k < dim; k++) inv[k] = 1.0 / box[k];
    double *xs = malloc(sizeof(double) * n * dim);
    long double *F = calloc(n * dim, sizeof(long double));
    if (!xs || !F) { free(xs); free(F); return -1; }
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < dim; k++) xs[i * dim + k] = wrap_coord(x[i * dim + k], box[k], inv[k]);
    if (nbr) memset(nbr, 0, sizeof(int32_t) * n);
    pair_acc acc = {0, 0, 0, 0};
    double cutoff2 = cutoff * cutoff;
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = i + 1; j < n; j++) {
            double r[3], d2;
            if (min_image(dim, xs + i * dim, xs + j * dim, box, inv, cutoff2, r, &d2))
                pair_update(dim, r, d2, i, j, diam, tag, p, F, &acc, nbr);
        }
    finish(dim, n, F, &acc, Fout, E, W, n_cut, n_int);
    free(xs); free(F);
    return 0;
}

/* Cell-list enumeration (linked cells, half shell), same arithmetic as brute force.
 * Falls back to brute force when a dimension has fewer than 3 cells. */
ORC_API int orc_forces(int dim, int64_t n, const double *x, const double *diam, const double *box, double cutoff,
                       int tag, const double *p, double *Fout, double *E, double *W, int64_t *n_cut, int64_t *n_int,
                       int32_t *nbr)
{
    int nc[3] = {1, 1, 1};
    double inv[3] = {1, 1, 1};
    for (int k = 0; k < dim; k++) {
        inv[k] = 1.0 / box[k];
        nc[k] = (int)floor(box[k] / (cutoff * (1.0 + 1e-9)));
        if (nc[k] < 3) return orc_forces_brute(dim, n, x, diam, box, cutoff, tag, p, Fout, E, W, n_cut, n_int, nbr);
    }
    int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
    double *xs = malloc(sizeof(double) * n * dim);
    long double *F = calloc(n * dim, sizeof(long double));
    int64_t *head = malloc(sizeof(int64_t) * ncell);
    int64_t *next = malloc(sizeof(int64_t) * n);
    if (!xs || !F || !head || !next) { free(xs); free(F); free(head); free(next); return -1; }
    for (int64_t c = 0; c < ncell; c++) head[c] = -1;
    if (nbr) memset(nbr, 0, sizeof(int32_t) * n);
    for (int64_t i = n - 1; i >= 0; i--) {
        int cc[3] = {0, 0, 0};
        for (int k = 0; k < dim; k++) {
            double v = wrap_coord(x[i * dim + k], box[k], inv[k]);
            xs[i * dim + k] = v;
            int c = (int)(v * inv[k] * nc[k]);
            if (c >= nc[k]) c = nc[k] - 1;
            if (c < 0) c = 0;
            cc[k] = c;
        }
        int64_t c = ((int64_t)cc[2] * nc[1] + cc[1]) * nc[0] + cc[0];
        next[i] = head[c];
        head[c] = i;
    }
    pair_acc acc = {0, 0, 0, 0};
    double cutoff2 = cutoff * cutoff;
    int zlo = (dim == 3) ? -1 : 0, zhi = (dim == 3) ? 1 : 0;
    for (int cz = 0; cz < nc[2]; cz++)
    for (int cy = 0; cy < nc[1]; cy++)
    for (int cx = 0; cx < nc[0]; cx++) {
        int64_t c = ((int64_t)cz * nc[1] + cy) * nc[0] + cx;
        for (int64_t i = head[c]; i >= 0; i = next[i]) {
            double r[3], d2;
            for (int64_t j = next[i]; j >= 0; j = next[j]) {
                int64_t a = i < j ? i : j, b = i < j ? j : i;
                if (min_image(dim, xs + a * dim, xs + b * dim, box, inv, cutoff2, r, &d2))
                    pair_update(dim, r, d2, a, b, diam, tag, p, F, &acc, nbr);
            }
        }
        for (int dz = zlo; dz <= zhi; dz++)
        for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            /* forward half shell: lexicographically positive (dz,dy,dx) */
            if (dz < 0 || (dz == 0 && (dy < 0 || (dy == 0 && dx <= 0)))) continue;
            int ox = (cx + dx + nc[0]) % nc[0], oy = (cy + dy + nc[1]) % nc[1], oz = (cz + dz + nc[2]) % nc[2];
            int64_t o = ((int64_t)oz * nc[1] + oy) * nc[0] + ox;
            for (int64_t i = head[c]; i >= 0; i = next[i])
                for (int64_t j = head[o]; j >= 0; j = next[j]) {
                    double r[3], d2;
                    /* keep r = x_lo - x_hi (lower original index first), like the i<j uniqueness of the
                       reference's unordered pairs; the sign convention cancels in every accumulated term */
                    int64_t a = i < j ? i : j, b = i < j ? j : i;
                    if (min_image(dim, xs + a * dim, xs + b * dim, box, inv, cutoff2, r, &d2))
                        pair_update(dim, r, d2, a, b, diam, tag, p, F, &acc, nbr);
                }
        }
    }
    finish(dim, n, F, &acc, Fout, E, W, n_cut, n_int);
    free(xs); free(F); free(head); free(next);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Boundary + integrators
 * ---------------------------------------------------------------------------------------- */

/* src/boundary.jl:7-17 wrap_to_box for a diagonal unit cell: frac = x/L (as inv(U)*x with
 * inv(U)=diag(1/L)), n = floor(frac), image += Int(n), x = L*(frac-n). */
ORC_API void orc_wrap(int dim, double *x, int32_t *img, const double *box)
{
    for (int k = 0; k < dim; k++) {
        double invL = 1.0 / box[k];
        double frac = invL * x[k];
        double nc = floor(frac);
        double fm = frac - nc;
        img[k] += (int32_t)nc;
        x[k] = box[k] * fm;
    }
}

/* src/integrate.jl:8-21 integrate_half!: v += (f*dt)/2; x += v*dt; x = wrap_to_box(x, img) */
ORC_API void orc_integrate_half(int dim, int64_t n, double *x, int32_t *img, double *v, const double *f, double dt,
                                const double *box)
{
    for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < dim; k++) {
            v[i * dim + k] += f[i * dim + k] * dt / 2.0;
            x[i * dim + k] += v[i * dim + k] * dt;
        }
        orc_wrap(dim, x + i * dim, img + i * dim, box);
    }
}

/* src/integrate.jl:28-38 integrate_second_half! */
ORC_API void orc_integrate_second_half(int dim, int64_t n, double *v, const double *f, double dt)
{
    for (int64_t q = 0; q < n * dim; q++) v[q] += f[q] * dt / 2.0;
}

/* src/thermostat.jl:50-60 compute_kinetic: serial, index order, sum(abs2, v_i) folded left. */
ORC_API double orc_kinetic(int dim, int64_t n, const double *v)
{
    double ke = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double s = v[i * dim] * v[i * dim];
        for (int k = 1; k < dim; k++) s += v[i * dim + k] * v[i * dim + k];
        ke += s;
    }
    return ke / 2.0;
}

/* src/thermostat.jl:62-67 */
ORC_API double orc_temperature(int dim, int64_t n, const double *v, double nf) { return 2.0 * orc_kinetic(dim, n, v) / nf; }

/* src/thermostat.jl:20-48 bussi!: the scale factor from (KE, T0, nf, dt, tau) and the two noises. */
ORC_API double orc_bussi_scale(double kinetic_energy, double ktemp, double nf, double dt, double tau, double r1, double r2)
{
    double dt_ratio = dt / tau;
    double current_temperature = 2.0 * kinetic_energy / nf;
    double term_1 = exp(-dt_ratio);
    double c2 = (1.0 - term_1) * ktemp / (current_temperature * nf);
    double term_2 = c2 * (r2 + r1 * r1);
    double term_3 = 2.0 * r1 * sqrt(term_1 * c2);
    return sqrt(term_1 + term_2 + term_3);
}

/* ------------------------------------------------------------------------------------------
 * Counter-based RNG.  The reference draws from Random.Xoshiro + Distributions.Gamma
 * (src/thermostat.jl:1-18,35-36; src/integrate.jl:55-59), neither reproducible here; north_star
 * prescribes a counter-based generator instead.  Spec (shared with csrc/rng.cuh):
 *   Philox4x32-10, key = (seed_lo, seed_hi);
 *   u53(w_hi, w_lo)      = ((w_hi<<32 | w_lo) >> 11) * 2^-53            in [0,1)
 *   u53_open(w_hi, w_lo) = (((w_hi<<32 | w_lo) >> 11) + 1) * 2^-53      in (0,1]
 *   thermostat normals : ctr = (step_lo, step_hi, block, 0xB0551<<8|0), Box-Muller on
 *                        (u53_open(w0,w1), u53(w2,w3)) -> n0 = r cos(2 pi u2), n1 = r sin(2 pi u2)
 *   thermostat uniforms: ctr = (step_lo, step_hi, block, 0xB0551<<8|1), two u53_open per block
 *   Brownian noise     : ctr = (particle_id, step_lo, step_hi, 0xB12D<<8|block); block 0 gives
 *                        u_x=(w0,w1), u_y=(w2,w3); block 1 gives u_z=(w0,w1)
 *   velocity normals   : ctr = (particle_id, stream_lo, stream_hi, 0x1E10C<<8|block), Box-Muller as above; block 0
 *                        gives (v_x, v_y), block 1 gives v_z = r cos (orc_init_velocities)
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        if (r > 0) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double u53(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1.0p-53; }
static inline double u53_open(uint32_t hi, uint32_t lo) { return (double)(((((uint64_t)hi << 32) | lo) >> 11) + 1) * 0x1.0p-53; }

#define ORC_TAG_BUSSI 0xB0551u
#define ORC_TAG_BROWN 0xB12Du

typedef struct {
    uint32_t key[2];
    uint64_t step;
    uint32_t nblock, ublock; /* next normal / uniform block */
    int nhave, uhave;
    double nbuf[2], ubuf[2];
} thermo_rng;

static void trng_init(thermo_rng *g, uint64_t seed, uint64_t step)
{
    g->key[0] = (uint32_t)seed; g->key[1] = (uint32_t)(seed >> 32);
    g->step = step; g->nblock = 0; g->ublock = 0; g->nhave = 0; g->uhave = 0;
}

static double trng_normal(thermo_rng *g)
{
    if (g->nhave == 0) {
        uint32_t ctr[4] = {(uint32_t)g->step, (uint32_t)(g->step >> 32), g->nblock++, (ORC_TAG_BUSSI << 8) | 0u}, w[4];
        orc_philox4x32_10(ctr, g->key, w);
        double u1 = u53_open(w[0], w[1]), u2 = u53(w[2], w[3]);
        double r = sqrt(-2.0 * log(u1));
        double th = 6.283185307179586 * u2;
        g->nbuf[0] = r * cos(th);
        g->nbuf[1] = r * sin(th);
        g->nhave = 2;
    }
    double v = g->nbuf[2 - g->nhave];
    g->nhave--;
    return v;
}

static double trng_uniform(thermo_rng *g)
{
    if (g->uhave == 0) {
        uint32_t ctr[4] = {(uint32_t)g->step, (uint32_t)(g->step >> 32), g->ublock++, (ORC_TAG_BUSSI << 8) | 1u}, w[4];
        orc_philox4x32_10(ctr, g->key, w);
        g->ubuf[0] = u53_open(w[0], w[1]);
        g->ubuf[1] = u53_open(w[2], w[3]);
        g->uhave = 2;
    }
    double v = g->ubuf[2 - g->uhave];
    g->uhave--;
    return v;
}

/* Gamma(k,1), k >= 1, Marsaglia-Tsang (the algorithm behind Distributions.Gamma's sampler for shape >= 1) */
static double trng_gamma(thermo_rng *g, double k)
{
    double d = k - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        double x = trng_normal(g);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        double u = trng_uniform(g);
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return d * v;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return d * v;
    }
}

/* src/thermostat.jl:1-18 sum_noises(nf, rng): chi^2 with nf degrees of freedom */
static double trng_sum_noises(thermo_rng *g, double nf)
{
    if (nf == 0.0) return 0.0;
    if (nf == 1.0) { double z = trng_normal(g); return z * z; }
    if (fmod(nf, 2.0) == 0.0) return 2.0 * trng_gamma(g, floor(nf / 2.0));
    double r = 2.0 * trng_gamma(g, floor((nf - 1.0) / 2.0));
    double z = trng_normal(g);
    return r + z * z;
}

/* the (r1, r2) pair bussi! draws at src/thermostat.jl:35-36: r1 = randn, r2 = sum_noises(nf - 1) */
ORC_API void orc_bussi_noises(uint64_t seed, uint64_t step, double nf, double *r1, double *r2)
{
    thermo_rng g;
    trng_init(&g, seed, step);
    *r1 = trng_normal(&g);
    *r2 = trng_sum_noises(&g, nf - 1.0);
}

/* raw streams, for distribution tests */
ORC_API void orc_thermo_normals(uint64_t seed, uint64_t step, int64_t count, double *out)
{
    thermo_rng g;
    trng_init(&g, seed, step);
    for (int64_t i = 0; i < count; i++) out[i] = trng_normal(&g);
}

ORC_API double orc_chi2(uint64_t seed, uint64_t step, double nf)
{
    thermo_rng g;
    trng_init(&g, seed, step);
    return trng_sum_noises(&g, nf);
}

/* src/integrate.jl:55-59 sample_uniform!: (2u-1)*sqrt(3), u ~ U[0,1) */
ORC_API void orc_brownian_noise(uint64_t seed, uint64_t step, uint32_t id, int dim, double *noise)
{
    const double sqthree = sqrt(3.0);
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, w[4];
    uint32_t ctr[4] = {id, (uint32_t)step, (uint32_t)(step >> 32), (ORC_TAG_BROWN << 8) | 0u};
    orc_philox4x32_10(ctr, key, w);
    noise[0] = (2.0 * u53(w[0], w[1]) - 1.0) * sqthree;
    noise[1] = (2.0 * u53(w[2], w[3]) - 1.0) * sqthree;
    if (dim == 3) {
        ctr[3] = (ORC_TAG_BROWN << 8) | 1u;
        orc_philox4x32_10(ctr, key, w);
        noise[2] = (2.0 * u53(w[0], w[1]) - 1.0) * sqthree;
    }
}

/* src/initialization.jl:22-27 initialize_random, first half: rand(rng, dim) .* (maxs .- mins) .+ mins with mins = 0.
 *   position uniforms: ctr = (particle_id, stream_lo, stream_hi, 0x9051<<8 | block); block 0 gives u_x = (w0,w1),
 *   u_y = (w2,w3); block 1 gives u_z = (w0,w1); x_k = L_k * u53, or 0 if the product rounds up to L_k. */
#define ORC_TAG_POS 0x9051u
ORC_API void orc_random_positions(int dim, int64_t n, const double *box, uint64_t seed, uint64_t stream, double *x)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, w[4];
    for (int64_t i = 0; i < n; i++) {
        uint32_t ctr[4] = {(uint32_t)i, (uint32_t)stream, (uint32_t)(stream >> 32), (ORC_TAG_POS << 8) | 0u};
        orc_philox4x32_10(ctr, key, w);
        x[i * dim + 0] = box[0] * u53(w[0], w[1]);
        x[i * dim + 1] = box[1] * u53(w[2], w[3]);
        if (dim == 3) {
            ctr[3] = (ORC_TAG_POS << 8) | 1u;
            orc_philox4x32_10(ctr, key, w);
            x[i * dim + 2] = box[2] * u53(w[0], w[1]);
        }
        for (int k = 0; k < dim; k++)
            if (!(x[i * dim + k] < box[k])) x[i * dim + k] = 0.0;
    }
}

/* src/initialization.jl:32-47 initialize_velocities: V = randn(d, N); V .-= mean(V; dims=2);
 * fs = sqrt(ktemp / (sum(abs2, V) / ((N-1) d))); V .*= fs.  The normals come from the counter-based generator:
 *   velocity normals: ctr = (particle_id, stream_lo, stream_hi, 0x1E10C<<8 | block), Box-Muller on
 *   (u53_open(w0,w1), u53(w2,w3)); block 0 gives (v_x, v_y) = (r cos, r sin), block 1 gives v_z = r cos.
 * v is [n][dim] in particle order; sums are accumulated in long double. */
#define ORC_TAG_VEL 0x1E10Cu
ORC_API void orc_init_velocities(int dim, int64_t n, double ktemp, uint64_t seed, uint64_t stream, double *v)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, w[4];
    long double mean[3] = {0, 0, 0};
    for (int64_t i = 0; i < n; i++) {
        uint32_t ctr[4] = {(uint32_t)i, (uint32_t)stream, (uint32_t)(stream >> 32), (ORC_TAG_VEL << 8) | 0u};
        orc_philox4x32_10(ctr, key, w);
        double r = sqrt(-2.0 * log(u53_open(w[0], w[1])));
        double th = 6.283185307179586 * u53(w[2], w[3]);
        v[i * dim + 0] = r * cos(th);
        v[i * dim + 1] = r * sin(th);
        if (dim == 3) {
            ctr[3] = (ORC_TAG_VEL << 8) | 1u;
            orc_philox4x32_10(ctr, key, w);
            double r2 = sqrt(-2.0 * log(u53_open(w[0], w[1])));
            v[i * dim + 2] = r2 * cos(6.283185307179586 * u53(w[2], w[3]));
        }
        for (int k = 0; k < dim; k++) mean[k] += v[i * dim + k];
    }
    double m[3] = {0, 0, 0};
    for (int k = 0; k < dim; k++) m[k] = (double)(mean[k] / (long double)n);
    long double sum_v2 = 0;
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < dim; k++) {
            v[i * dim + k] -= m[k];
            sum_v2 += (long double)v[i * dim + k] * v[i * dim + k];
        }
    double fs = sqrt(ktemp / ((double)sum_v2 / (((double)n - 1.0) * dim)));
    for (int64_t i = 0; i < n * dim; i++) v[i] *= fs;
}

/* src/integrate.jl:66-82 integrate_brownian! with the *intended* semantics:
 * x = x + (f*dt/kT) + (noise*sigma), sigma = sqrt(2 dt) (src/simulation.jl:212), then wrap_to_box.
 * ids[i] is the particle's original index (the RNG counter). */
ORC_API void orc_integrate_brownian(int dim, int64_t n, double *x, int32_t *img, const double *f, double dt, double ktemp,
                                    const double *box, uint64_t seed, uint64_t step, const int32_t *ids)
{
    double sigma = sqrt(2.0 * dt);
    for (int64_t i = 0; i < n; i++) {
        double noise[3];
        orc_brownian_noise(seed, step, ids ? (uint32_t)ids[i] : (uint32_t)i, dim, noise);
        for (int k = 0; k < dim; k++) x[i * dim + k] = x[i * dim + k] + (f[i * dim + k] * dt / ktemp) + (noise[k] * sigma);
        orc_wrap(dim, x + i * dim, img + i * dim, box);
    }
}

/* ------------------------------------------------------------------------------------------
 * Whole loops (src/simulation.jl:88-108 and :231-250).  thermo rows per step: [U, W, KE, n_int].
 * ensemble: 0 NVE, 1 NVT (ktemp_per_step[s] = ensemble.ktemp(step+1), src/integrate.jl:49), 2 Brownian.
 * `rng_step0` = engine-wide step counter at the start of the call (RNG counter base).
 * Forces are not primed before step 0: x, v, f are in/out.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_run(int ensemble, int dim, int64_t n, double *x, double *v, double *f, int32_t *img, const double *diam,
                    const double *box, double cutoff, int tag, const double *p, double dt, int64_t nsteps,
                    const double *ktemp_per_step, double tau, double nf, uint64_t seed, uint64_t rng_step0,
                    double *thermo /* [nsteps][4] or NULL */)
{
    for (int64_t s = 0; s < nsteps; s++) {
        double E, W;
        int64_t n_cut, n_int;
        double ke;
        if (ensemble == 2) {
            if (orc_forces(dim, n, x, diam, box, cutoff, tag, p, f, &E, &W, &n_cut, &n_int, NULL)) return -1;
            orc_integrate_brownian(dim, n, x, img, f, dt, ktemp_per_step[0], box, seed, rng_step0 + s, NULL);
            ke = 0.0;
        } else {
            orc_integrate_half(dim, n, x, img, v, f, dt, box);
            if (orc_forces(dim, n, x, diam, box, cutoff, tag, p, f, &E, &W, &n_cut, &n_int, NULL)) return -1;
            orc_integrate_second_half(dim, n, v, f, dt);
            if (ensemble == 1) {
                double r1, r2;
                double k0 = orc_kinetic(dim, n, v);
                orc_bussi_noises(seed, rng_step0 + s, nf, &r1, &r2);
                double scale = orc_bussi_scale(k0, ktemp_per_step[s], nf, dt, tau, r1, r2);
                for (int64_t q = 0; q < n * dim; q++) v[q] = v[q] * scale;
            }
            ke = orc_kinetic(dim, n, v);
        }
        if (thermo) {
            thermo[4 * s + 0] = E; thermo[4 * s + 1] = W; thermo[4 * s + 2] = ke; thermo[4 * s + 3] = (double)n_int;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * General (triclinic) unit cells.  The cell is a dim x dim matrix with the lattice vectors in its columns, x = U*frac
 * (to_unitcell, src/initialization.jl:7-18; wrap_to_box, src/boundary.jl:7-17; CellListMap accepts the same matrix,
 * src/initialization.jl:100-107).  U is passed row-major embedded in 3x3 (U[3*r+c]; a 2-D cell has U[8] = 1).
 * Matrix-vector products sum each row left to right.  Pairs: every unordered pair whose nearest periodic image,
 * d - U*nearbyint(U^-1 d), lies within the cutoff (valid for cutoff < half the smallest perpendicular width).
 * ---------------------------------------------------------------------------------------- */
ORC_API double orc_cell_inverse(const double *U, double *Ui)
{
    const double a = U[0], b = U[1], c = U[2], d = U[3], e = U[4], f = U[5], g = U[6], h = U[7], i = U[8];
    const double A = e * i - f * h, B = f * g - d * i, C = d * h - e * g;
    const double det = a * A + b * B + c * C;
    Ui[0] = A / det;             Ui[1] = (c * h - b * i) / det; Ui[2] = (b * f - c * e) / det;
    Ui[3] = B / det;             Ui[4] = (a * i - c * g) / det; Ui[5] = (c * d - a * f) / det;
    Ui[6] = C / det;             Ui[7] = (b * g - a * h) / det; Ui[8] = (a * e - b * d) / det;
    return det;
}

static inline void mat3_mul(const double *M, const double *v, double *o)
{
    for (int r = 0; r < 3; r++) o[r] = M[3 * r] * v[0] + M[3 * r + 1] * v[1] + M[3 * r + 2] * v[2];
}

/* src/boundary.jl:7-17 with a full matrix */
ORC_API void orc_wrap_tri(int dim, double *x, int32_t *img, const double *U, const double *Ui)
{
    double v[3] = {x[0], x[1], dim == 3 ? x[2] : 0.0}, fr[3], o[3];
    mat3_mul(Ui, v, fr);
    for (int k = 0; k < 3; k++) {
        double nc = k < dim ? floor(fr[k]) : 0.0;
        fr[k] = k < dim ? fr[k] - nc : 0.0;
        if (k < dim) img[k] += (int32_t)nc;
    }
    mat3_mul(U, fr, o);
    for (int k = 0; k < dim; k++) x[k] = o[k];
}

ORC_API int orc_forces_tri(int dim, int64_t n, const double *x, const double *diam, const double *U, double cutoff, int tag,
                           const double *p, double *Fout, double *E, double *W, int64_t *n_cut, int64_t *n_int)
{
    double Ui[9];
    orc_cell_inverse(U, Ui);
    double *xs = malloc(sizeof(double) * n * dim);
    long double *F = calloc(n * dim, sizeof(long double));
    if (!xs || !F) { free(xs); free(F); return -1; }
    for (int64_t i = 0; i < n; i++) {
        int32_t scratch[3] = {0, 0, 0};
        for (int k = 0; k < dim; k++) xs[i * dim + k] = x[i * dim + k];
        orc_wrap_tri(dim, xs + i * dim, scratch, U, Ui);
    }
    pair_acc acc = {0, 0, 0, 0};
    double cutoff2 = cutoff * cutoff;
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = i + 1; j < n; j++) {
            double d[3] = {0, 0, 0}, fr[3], kk[3], sh[3];
            for (int k = 0; k < dim; k++) d[k] = xs[i * dim + k] - xs[j * dim + k];
            mat3_mul(Ui, d, fr);
            for (int k = 0; k < 3; k++) kk[k] = k < dim ? nearbyint(fr[k]) : 0.0;
            mat3_mul(U, kk, sh);
            double d2 = 0.0;
            for (int k = 0; k < dim; k++) {
                d[k] = d[k] - sh[k];
                d2 = (k == 0) ? d[k] * d[k] : fma(d[k], d[k], d2);
            }
            if (d2 <= cutoff2) pair_update(dim, d, d2, i, j, diam, tag, p, F, &acc, NULL);
        }
    finish(dim, n, F, &acc, Fout, E, W, n_cut, n_int);
    free(xs); free(F);
    return 0;
}

/* the loops of orc_run with the general cell (all-pairs enumeration: small systems) */
ORC_API int orc_run_tri(int ensemble, int dim, int64_t n, double *x, double *v, double *f, int32_t *img, const double *diam,
                        const double *U, double cutoff, int tag, const double *p, double dt, int64_t nsteps,
                        const double *ktemp_per_step, double tau, double nf, uint64_t seed, uint64_t rng_step0, double *thermo)
{
    double Ui[9];
    orc_cell_inverse(U, Ui);
    for (int64_t s = 0; s < nsteps; s++) {
        double E, W, ke = 0.0;
        int64_t n_cut, n_int;
        if (ensemble == 2) {
            if (orc_forces_tri(dim, n, x, diam, U, cutoff, tag, p, f, &E, &W, &n_cut, &n_int)) return -1;
            double sigma = sqrt(2.0 * dt);
            for (int64_t i = 0; i < n; i++) {
                double noise[3];
                orc_brownian_noise(seed, rng_step0 + s, (uint32_t)i, dim, noise);
                for (int k = 0; k < dim; k++)
                    x[i * dim + k] = x[i * dim + k] + (f[i * dim + k] * dt / ktemp_per_step[0]) + (noise[k] * sigma);
                orc_wrap_tri(dim, x + i * dim, img + i * dim, U, Ui);
            }
        } else {
            for (int64_t i = 0; i < n; i++) {
                for (int k = 0; k < dim; k++) {
                    v[i * dim + k] += f[i * dim + k] * dt / 2.0;
                    x[i * dim + k] += v[i * dim + k] * dt;
                }
                orc_wrap_tri(dim, x + i * dim, img + i * dim, U, Ui);
            }
            if (orc_forces_tri(dim, n, x, diam, U, cutoff, tag, p, f, &E, &W, &n_cut, &n_int)) return -1;
            orc_integrate_second_half(dim, n, v, f, dt);
            if (ensemble == 1) {
                double r1, r2;
                double k0 = orc_kinetic(dim, n, v);
                orc_bussi_noises(seed, rng_step0 + s, nf, &r1, &r2);
                double scale = orc_bussi_scale(k0, ktemp_per_step[s], nf, dt, tau, r1, r2);
                for (int64_t q = 0; q < n * dim; q++) v[q] = v[q] * scale;
            }
            ke = orc_kinetic(dim, n, v);
        }
        if (thermo) {
            thermo[4 * s + 0] = E; thermo[4 * s + 1] = W; thermo[4 * s + 2] = ke; thermo[4 * s + 3] = (double)n_int;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Timing variant: "reference-equivalent CPU path" (BASELINE.md section 3).  Same algorithmic shape as
 * the Julia package: cell list rebuilt EVERY step with cell = cutoff (src/simulation.jl:100-104),
 * one visit per unordered pair with Newton's third law (src/pairwise.jl:35-36), per-thread
 * private force arrays merged afterwards (src/pairwise.jl:17-23), separate threaded integrate sweeps
 * (src/integrate.jl:8-38), serial kinetic energy (src/thermostat.jl:50-60).  Plain double
 * accumulation.  Returns the number of steps done; E/W/KE of the last step in out[3].
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the timing baseline asks for the host's cores explicitly */
ORC_API void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_API int orc_run_timing(int ensemble, int dim, int64_t n, double *x, double *v, double *f, int32_t *img,
                           const double *diam, const double *box, double cutoff, int tag, const double *p, double dt,
                           int64_t nsteps, double ktemp, double tau, double nf, uint64_t seed, double *out)
{
    int nc[3] = {1, 1, 1};
    double inv[3] = {1, 1, 1};
    for (int k = 0; k < dim; k++) {
        inv[k] = 1.0 / box[k];
        nc[k] = (int)floor(box[k] / (cutoff * (1.0 + 1e-9)));
        if (nc[k] < 3) return -2;
    }
    int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
    int nth = orc_threads();
    int64_t *start = malloc(sizeof(int64_t) * (ncell + 1));
    int64_t *cellof = malloc(sizeof(int64_t) * n);
    int64_t *order = malloc(sizeof(int64_t) * n);
    double *xs = malloc(sizeof(double) * n * dim);
    double *fth = malloc(sizeof(double) * (size_t)nth * n * dim);
    double *eth = calloc((size_t)nth * 8, sizeof(double));
    if (!start || !cellof || !order || !xs || !fth || !eth) return -1;
    double cutoff2 = cutoff * cutoff;
    int zlo = (dim == 3) ? -1 : 0, zhi = (dim == 3) ? 1 : 0;
    double E = 0, W = 0, ke = 0;
    for (int64_t s = 0; s < nsteps; s++) {
        if (ensemble != 2) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; i++) {
                for (int k = 0; k < dim; k++) {
                    v[i * dim + k] += f[i * dim + k] * dt / 2.0;
                    x[i * dim + k] += v[i * dim + k] * dt;
                }
                orc_wrap(dim, x + i * dim, img + i * dim, box);
            }
        }
        /* cell list rebuild (counting sort) */
        memset(start, 0, sizeof(int64_t) * (ncell + 1));
        for (int64_t i = 0; i < n; i++) {
            int cc[3] = {0, 0, 0};
            for (int k = 0; k < dim; k++) {
                double w = wrap_coord(x[i * dim + k], box[k], inv[k]);
                xs[i * dim + k] = w;
                int c = (int)(w * inv[k] * nc[k]);
                if (c >= nc[k]) c = nc[k] - 1;
                cc[k] = c;
            }
            cellof[i] = ((int64_t)cc[2] * nc[1] + cc[1]) * nc[0] + cc[0];
            start[cellof[i] + 1]++;
        }
        for (int64_t c = 0; c < ncell; c++) start[c + 1] += start[c];
        {
            int64_t *fill = calloc(ncell, sizeof(int64_t));
            for (int64_t i = 0; i < n; i++) order[start[cellof[i]] + fill[cellof[i]]++] = i;
            free(fill);
        }
        /* reset_output! on every per-thread copy, then the pair map */
#pragma omp parallel
        {
#ifdef _OPENMP
            int t = omp_get_thread_num();
#else
            int t = 0;
#endif
            double *F = fth + (size_t)t * n * dim;
            memset(F, 0, sizeof(double) * n * dim);
            double e = 0, w = 0;
#pragma omp for schedule(dynamic, 64)
            for (int64_t c = 0; c < ncell; c++) {
                int cx = (int)(c % nc[0]), cy = (int)((c / nc[0]) % nc[1]), cz = (int)(c / ((int64_t)nc[0] * nc[1]));
                for (int dz = zlo; dz <= zhi; dz++)
                for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    if (dz < 0 || (dz == 0 && (dy < 0 || (dy == 0 && dx < 0)))) continue;
                    int self = (dz == 0 && dy == 0 && dx == 0);
                    int ox = (cx + dx + nc[0]) % nc[0], oy = (cy + dy + nc[1]) % nc[1], oz = (cz + dz + nc[2]) % nc[2];
                    int64_t o = ((int64_t)oz * nc[1] + oy) * nc[0] + ox;
                    for (int64_t a = start[c]; a < start[c + 1]; a++) {
                        int64_t i = order[a];
                        for (int64_t b = self ? a + 1 : start[o]; b < start[o + 1]; b++) {
                            int64_t j = order[b];
                            double r[3], d2;
                            if (!min_image(dim, xs + i * dim, xs + j * dim, box, inv, cutoff2, r, &d2)) continue;
                            double d = sqrt(d2), u, ff;
                            orc_evaluate(tag, p, d, diam[i], diam[j], &u, &ff);
                            double dot = 0;
                            for (int k = 0; k < dim; k++) {
                                double sk = (ff * r[k]) / d;
                                dot += sk * r[k];
                                F[i * dim + k] += sk;
                                F[j * dim + k] -= sk;
                            }
                            w += dot; e += u;
                        }
                    }
                }
            }
            eth[t * 8] = e; eth[t * 8 + 1] = w;
        }
        /* reducer: merge the per-thread copies */
#pragma omp parallel for schedule(static)
        for (int64_t q = 0; q < n * dim; q++) {
            double acc = 0;
            for (int t = 0; t < nth; t++) acc += fth[(size_t)t * n * dim + q];
            f[q] = acc;
        }
        E = 0; W = 0;
        for (int t = 0; t < nth; t++) { E += eth[t * 8]; W += eth[t * 8 + 1]; }
        if (ensemble == 2) {
            orc_integrate_brownian(dim, n, x, img, f, dt, ktemp, box, seed, (uint64_t)s, NULL);
        } else {
#pragma omp parallel for schedule(static)
            for (int64_t q = 0; q < n * dim; q++) v[q] += f[q] * dt / 2.0;
            if (ensemble == 1) {
                double r1, r2;
                double k0 = orc_kinetic(dim, n, v);
                orc_bussi_noises(seed, (uint64_t)s, nf, &r1, &r2);
                double scale = orc_bussi_scale(k0, ktemp, nf, dt, tau, r1, r2);
                for (int64_t q = 0; q < n * dim; q++) v[q] = v[q] * scale;
            }
            ke = orc_kinetic(dim, n, v);
        }
    }
    if (out) { out[0] = E; out[1] = W; out[2] = ke; }
    free(start); free(cellof); free(order); free(xs); free(fth); free(eth);
    return (int)nsteps;
}

/* ------------------------------------------------------------------------------------------
 * FIRE minimiser: src/minimize.jl:31-135 fire_minimize!, statement by statement.
 * x, img in/out.  trace (optional, [max_steps][3]) receives per step {energy, F_norm/sqrt(ndof), dt used for the move}.
 * Returns the number of steps performed (force evaluations in the loop); *converged = 1 when the F_rms test passed.
 * ---------------------------------------------------------------------------------------- */
ORC_API int64_t orc_fire(int dim, int64_t n, double *x, int32_t *img, const double *diam, const double *box, double cutoff,
                         int tag, const double *p, int64_t max_steps, double tol, double dt_initial, double dt_max,
                         double alpha0, double f_inc, double f_dec, int n_min, double *energy_out, int *converged,
                         double *trace)
{
    double alpha = alpha0;
    int64_t steps_since_neg = 0;
    double dt = dt_initial;
    double *v = calloc(n * dim, sizeof(double));
    double *f = calloc(n * dim, sizeof(double));
    double ndof = dim * (n - 1.0);
    double E = 0.0, W;
    *converged = 0;
    int64_t step;
    for (step = 1; step <= max_steps; step++) {
        orc_forces(dim, n, x, diam, box, cutoff, tag, p, f, &E, &W, NULL, NULL, NULL);
        double ff = 0.0;
        for (int64_t q = 0; q < n * dim; q++) ff += f[q] * f[q];
        double F_norm = sqrt(ff);
        if (trace) { trace[3 * (step - 1)] = E; trace[3 * (step - 1) + 1] = F_norm / sqrt(ndof); trace[3 * (step - 1) + 2] = 0.0; }
        if (F_norm / sqrt(ndof) < tol) {
            *converged = 1;
            break;
        }
        for (int64_t q = 0; q < n * dim; q++) v[q] += dt * f[q];
        double P = 0.0, vv = 0.0;
        for (int64_t q = 0; q < n * dim; q++) { P += v[q] * f[q]; vv += v[q] * v[q]; }
        double v_norm = sqrt(vv), f_norm = sqrt(ff);
        if (v_norm > 0 && f_norm > 0) {
            double scale = alpha * (v_norm / f_norm);
            for (int64_t q = 0; q < n * dim; q++) v[q] = (1.0 - alpha) * v[q] + scale * f[q];
        }
        if (P > 0) {
            steps_since_neg += 1;
            if (steps_since_neg > n_min) {
                dt = fmin(dt * f_inc, dt_max);
                alpha *= 0.99;
            }
        } else {
            dt = fmax(dt * f_dec, dt_initial);
            memset(v, 0, sizeof(double) * n * dim);
            alpha = alpha0;
            steps_since_neg = 0;
        }
        if (trace) trace[3 * (step - 1) + 2] = dt;
        for (int64_t i = 0; i < n; i++) {
            for (int k = 0; k < dim; k++) x[i * dim + k] += dt * v[i * dim + k];
            orc_wrap(dim, x + i * dim, img + i * dim, box);
        }
    }
    if (!*converged) {
        step = max_steps;
        orc_forces(dim, n, x, diam, box, cutoff, tag, p, f, &E, &W, NULL, NULL, NULL);
    }
    *energy_out = E;
    free(v); free(f);
    return step;
}
