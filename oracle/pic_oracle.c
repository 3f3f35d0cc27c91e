/*
 * pic_oracle.c -- CPU restatement of the lambdaPIC per-step inner loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (lambdapic_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 *   (1) the reference's own C extensions compiled into oracle/_ref/ (same inputs, call by call), and
 *   (2) tests/golden/ref_step_{2d,3d}.npz, produced by running the unmodified reference
 *       (oracle/make_golden.py).
 * Integer results (sort permutation, bucket tables, migration slots, capacities, is_dead) are
 * bit-exact; floating-point results are bit-exact for FDTD / guard copy / current reduce and agree
 * to <= 1e-13 relative for gather+Boris+deposit (the reference is built with gcc FMA contraction,
 * this file with -ffp-contract=off).
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference/src/lambdapic/).  Layout conventions are the reference's:
 *   - a field grid is C-contiguous (NX,NY,NZ) = (nx+2ng, ny+2ng, nz+2ng); logical index i in
 *     [-ng, n+ng) lives at storage index (i < 0 ? i + N : i)      (core/fields.py:24-26, cutils.h:19-26)
 *   - 2D grids are the same with the z axis absent (NZ = 1, no z guards).
 *   - particles are SoA fp64 arrays + uint8 is_dead                (core/particles.py:63-67)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define C_LIGHT 299792458.0          /* cutils.h:17 */
#define ONE_THIRD 0.3333333333333333 /* cutils.h:18 */

typedef int64_t i64;
typedef uint8_t u8;

static inline i64 wrapneg(i64 i, i64 N) { return i >= 0 ? i : i + N; }

/* ------------------------------------------------------------------------------------------------
 * Boundary tables.  Enum order = core/patch/sync_fields3d.c:19-50 (3D) and sync_fields2d.c:18-28 (2D).
 * dir[b] = (sx,sy,sz) with -1 = MIN side, +1 = MAX side.
 * ---------------------------------------------------------------------------------------------- */
static const int DIR3[26][3] = {
    {-1,0,0},{1,0,0},{0,-1,0},{0,1,0},{0,0,-1},{0,0,1},
    {-1,-1,0},{-1,1,0},{-1,0,-1},{-1,0,1},{1,-1,0},{1,1,0},{1,0,-1},{1,0,1},
    {0,-1,-1},{0,-1,1},{0,1,-1},{0,1,1},
    {-1,-1,-1},{-1,-1,1},{-1,1,-1},{-1,1,1},{1,-1,-1},{1,-1,1},{1,1,-1},{1,1,1}};
static const int DIR2[8][3] = {
    {-1,0,0},{1,0,0},{0,-1,0},{0,1,0},{-1,-1,0},{1,-1,0},{-1,1,0},{1,1,0}};

static int find_dir(const int (*tab)[3], int nb, int sx, int sy, int sz) {
    for (int b = 0; b < nb; b++)
        if (tab[b][0] == sx && tab[b][1] == sy && tab[b][2] == sz) return b;
    return -1;
}
static int opposite(const int (*tab)[3], int nb, int b) {
    return find_dir(tab, nb, -tab[b][0], -tab[b][1], -tab[b][2]);
}

/* ------------------------------------------------------------------------------------------------
 * Yee FDTD.  core/maxwell/cpu.py:83-97 (E, 3D), :101-112 (B, 3D), :9-35 (2D).
 * Expression order is the reference's (numba does not contract; this file is built without
 * contraction), so results are bit-identical.  bfactor = dt*c**2, jfactor = dt/epsilon_0 are
 * computed by the caller exactly as the reference does (cpu.py:90-91).
 * ---------------------------------------------------------------------------------------------- */
#define IX3(i, j, k) (wrapneg(k, NZ) + wrapneg(j, NY) * NZ + wrapneg(i, NX) * NY * NZ)

void orc_update_efield_3d(double *ex, double *ey, double *ez, const double *bx, const double *by, const double *bz,
                          const double *jx, const double *jy, const double *jz,
                          i64 nx, i64 ny, i64 nz, i64 ng, double dx, double dy, double dz,
                          double bfactor, double jfactor) {
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng, NZ = nz + 2 * ng;
    for (i64 i = 0; i < nx; i++)
        for (i64 j = 0; j < ny; j++)
            for (i64 k = 0; k < nz; k++) {
                i64 c = IX3(i, j, k), xm = IX3(i - 1, j, k), ym = IX3(i, j - 1, k), zm = IX3(i, j, k - 1);
                ex[c] += bfactor * ((bz[c] - bz[ym]) / dy - (by[c] - by[zm]) / dz) - jfactor * jx[c];
                ey[c] += bfactor * ((bx[c] - bx[zm]) / dz - (bz[c] - bz[xm]) / dx) - jfactor * jy[c];
                ez[c] += bfactor * ((by[c] - by[xm]) / dx - (bx[c] - bx[ym]) / dy) - jfactor * jz[c];
            }
}

void orc_update_bfield_3d(const double *ex, const double *ey, const double *ez, double *bx, double *by, double *bz,
                          i64 nx, i64 ny, i64 nz, i64 ng, double dx, double dy, double dz, double dt) {
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng, NZ = nz + 2 * ng;
    for (i64 i = 0; i < nx; i++)
        for (i64 j = 0; j < ny; j++)
            for (i64 k = 0; k < nz; k++) {
                i64 c = IX3(i, j, k), xp = IX3(i + 1, j, k), yp = IX3(i, j + 1, k), zp = IX3(i, j, k + 1);
                bx[c] -= dt * ((ez[yp] - ez[c]) / dy - (ey[zp] - ey[c]) / dz);
                by[c] -= dt * ((ex[zp] - ex[c]) / dz - (ez[xp] - ez[c]) / dx);
                bz[c] -= dt * ((ey[xp] - ey[c]) / dx - (ex[yp] - ex[c]) / dy);
            }
}

/* All patches of a rank in one call, OpenMP over patches as the reference's numba kernels run them
 * (update_efield_patches_3d / update_bfield_patches_3d, core/maxwell/cpu.py:37-158: prange over patches).
 * f = npatch x 9 (E) or npatch x 6 (B) array pointers in FIELD_ATTRS order; same per-patch arithmetic as above. */
void orc_update_efield_3d_all(double **f, i64 npatch, i64 nx, i64 ny, i64 nz, i64 ng, double dx, double dy, double dz,
                              double bfactor, double jfactor) {
#pragma omp parallel for schedule(static)
    for (i64 p = 0; p < npatch; p++) {
        double **a = f + 9 * p;
        orc_update_efield_3d(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], nx, ny, nz, ng, dx, dy, dz, bfactor, jfactor);
    }
}
void orc_update_bfield_3d_all(double **f, i64 npatch, i64 nx, i64 ny, i64 nz, i64 ng, double dx, double dy, double dz, double dt) {
#pragma omp parallel for schedule(static)
    for (i64 p = 0; p < npatch; p++) {
        double **a = f + 6 * p;
        orc_update_bfield_3d(a[0], a[1], a[2], a[3], a[4], a[5], nx, ny, nz, ng, dx, dy, dz, dt);
    }
}

#define IX2(i, j) (wrapneg(j, NY) + wrapneg(i, NX) * NY)

void orc_update_efield_2d(double *ex, double *ey, double *ez, const double *bx, const double *by, const double *bz,
                          const double *jx, const double *jy, const double *jz,
                          i64 nx, i64 ny, i64 ng, double dx, double dy, double bfactor, double jfactor) {
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng;
    for (i64 i = 0; i < nx; i++)
        for (i64 j = 0; j < ny; j++) {
            i64 c = IX2(i, j), xm = IX2(i - 1, j), ym = IX2(i, j - 1);
            ex[c] += bfactor * ((bz[c] - bz[ym]) / dy) - jfactor * jx[c];
            ey[c] += bfactor * (-(bz[c] - bz[xm]) / dx) - jfactor * jy[c];
            ez[c] += bfactor * ((by[c] - by[xm]) / dx - (bx[c] - bx[ym]) / dy) - jfactor * jz[c];
        }
}

void orc_update_bfield_2d(const double *ex, const double *ey, const double *ez, double *bx, double *by, double *bz,
                          i64 nx, i64 ny, i64 ng, double dx, double dy, double dt) {
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng;
    for (i64 i = 0; i < nx; i++)
        for (i64 j = 0; j < ny; j++) {
            i64 c = IX2(i, j), xp = IX2(i + 1, j), yp = IX2(i, j + 1);
            bx[c] -= dt * ((ez[yp] - ez[c]) / dy);
            by[c] -= dt * (-(ez[xp] - ez[c]) / dx);
            bz[c] -= dt * ((ey[xp] - ey[c]) / dx - (ex[yp] - ex[c]) / dy);
        }
}

/* ------------------------------------------------------------------------------------------------
 * Guard-cell copy (E/B) and guard->interior current reduce (J, rho).
 * core/patch/sync_fields3d.c:350-620 / :84-348, sync_fields2d.c:150-255 / :43-148.
 *
 * For a boundary with direction s along one axis (n interior cells, ng guards):
 *   guard copy   : s=-1: dst[-ng,0) <- src[n-ng,n);  s=+1: dst[n,n+ng) <- src[0,ng);  s=0: [0,n) <- [0,n)
 *   current sum  : s=-1: dst[0,ng) += src[n,n+ng);   s=+1: dst[n-ng,n) += src[-ng,0); s=0: [0,n) += [0,n)
 *                  and the source strip is zeroed.  Boundaries are visited in enum order, which fixes the
 *                  floating-point summation order at edge/corner cells.
 * `f` holds one pointer per patch for a single grid attribute; nbr is (npatch, nb) int64, <0 = none.
 * dim==2 uses the 8-entry table and ignores z.
 * ---------------------------------------------------------------------------------------------- */
static void axis_ranges(int s, i64 n, i64 ng, int reduce, i64 *dst0, i64 *src0, i64 *len) {
    if (s == 0) { *dst0 = 0; *src0 = 0; *len = n; }
    else if (!reduce) { if (s < 0) { *dst0 = -ng; *src0 = n - ng; } else { *dst0 = n; *src0 = 0; } *len = ng; }
    else { if (s < 0) { *dst0 = 0; *src0 = n; } else { *dst0 = n - ng; *src0 = -ng; } *len = ng; }
}

static void sync_generic(double **f, i64 npatch, const i64 *nbr, int dim, i64 nx, i64 ny, i64 nz, i64 ng, int reduce) {
    const int nb = dim == 3 ? 26 : 8;
    const int (*tab)[3] = dim == 3 ? DIR3 : DIR2;
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng, NZ = dim == 3 ? nz + 2 * ng : 1;
    if (dim == 2) nz = 1;
    for (i64 p = 0; p < npatch; p++)
        for (int b = 0; b < nb; b++) {
            i64 q = nbr[p * nb + b];
            if (q < 0) continue;
            i64 dx0, sx0, lx, dy0, sy0, ly, dz0 = 0, sz0 = 0, lz = 1;
            axis_ranges(tab[b][0], nx, ng, reduce, &dx0, &sx0, &lx);
            axis_ranges(tab[b][1], ny, ng, reduce, &dy0, &sy0, &ly);
            if (dim == 3) axis_ranges(tab[b][2], nz, ng, reduce, &dz0, &sz0, &lz);
            for (i64 i = 0; i < lx; i++)
                for (i64 j = 0; j < ly; j++)
                    for (i64 k = 0; k < lz; k++) {
                        i64 d = wrapneg(dz0 + k, NZ) + wrapneg(dy0 + j, NY) * NZ + wrapneg(dx0 + i, NX) * NY * NZ;
                        i64 s = wrapneg(sz0 + k, NZ) + wrapneg(sy0 + j, NY) * NZ + wrapneg(sx0 + i, NX) * NY * NZ;
                        if (reduce) { f[p][d] += f[q][s]; f[q][s] = 0.0; }
                        else f[p][d] = f[q][s];
                    }
        }
}
void orc_sync_guard(double **f, i64 npatch, const i64 *nbr, i64 dim, i64 nx, i64 ny, i64 nz, i64 ng) {
    sync_generic(f, npatch, nbr, (int)dim, nx, ny, nz, ng, 0);
}
void orc_sync_currents(double **f, i64 npatch, const i64 *nbr, i64 dim, i64 nx, i64 ny, i64 nz, i64 ng) {
    sync_generic(f, npatch, nbr, (int)dim, nx, ny, nz, ng, 1);
}

/* ------------------------------------------------------------------------------------------------
 * Gather (TSC / quadratic spline on the staggered Yee grid) + Boris + Esirkepov deposit.
 * core/pusher/unified/unified_pusher_3d.c:15-217 (boris, push, get_gx, interpolation_3d),
 * core/current/current_deposit.h:7-35,275-440 (S/S0, deposit_3d_fast), 2D: unified_pusher_2d.c:64-155,
 * current_deposit.h:185-268.
 * ---------------------------------------------------------------------------------------------- */
static inline void tsc3(double d, double *g) { /* get_gx, unified_pusher_3d.c:65-70 */
    double d2 = d * d;
    g[0] = 0.5 * (0.25 + d2 + d);
    g[1] = 0.75 - d2;
    g[2] = 0.5 * (0.25 + d2 - d);
}

static inline double gather27(const double *F, const double *fx, const double *fy, const double *fz,
                              i64 ix, i64 iy, i64 iz, i64 NX, i64 NY, i64 NZ) {
    /* nesting and association as interp_field_safe_3d, unified_pusher_3d.c:79-106 */
    double acc_z[3];
    for (int c = 0; c < 3; c++) {
        double acc_y[3];
        for (int b = 0; b < 3; b++) {
            i64 base = wrapneg(iz + c - 1, NZ) + wrapneg(iy + b - 1, NY) * NZ;
            acc_y[b] = fx[0] * F[base + wrapneg(ix - 1, NX) * NY * NZ] + fx[1] * F[base + wrapneg(ix, NX) * NY * NZ] +
                       fx[2] * F[base + wrapneg(ix + 1, NX) * NY * NZ];
        }
        acc_z[c] = fy[0] * acc_y[0] + fy[1] * acc_y[1] + fy[2] * acc_y[2];
    }
    return fz[0] * acc_z[0] + fz[1] * acc_z[1] + fz[2] * acc_z[2];
}

static inline double gather9(const double *F, const double *fx, const double *fy, i64 ix, i64 iy, i64 NX, i64 NY) {
    /* interp_field_safe, unified_pusher_2d.c:73-85 */
    double acc[3];
    for (int b = 0; b < 3; b++) {
        i64 base = wrapneg(iy + b - 1, NY);
        acc[b] = fx[0] * F[base + wrapneg(ix - 1, NX) * NY] + fx[1] * F[base + wrapneg(ix, NX) * NY] +
                 fx[2] * F[base + wrapneg(ix + 1, NX) * NY];
    }
    return fy[0] * acc[0] + fy[1] * acc[1] + fy[2] * acc[2];
}

static inline void boris_kick(double *ux, double *uy, double *uz, double *inv_gamma, double Ex, double Ey, double Ez,
                              double Bx, double By, double Bz, double efactor, double bfactor) {
    /* unified_pusher_3d.c:15-51 */
    double umx = *ux + efactor * Ex, umy = *uy + efactor * Ey, umz = *uz + efactor * Ez;
    double ig = 1.0 / sqrt(1 + umx * umx + umy * umy + umz * umz);
    double Tx = bfactor * Bx * ig, Ty = bfactor * By * ig, Tz = bfactor * Bz * ig;
    double upx = umx + umy * Tz - umz * Ty;
    double upy = umy + umz * Tx - umx * Tz;
    double upz = umz + umx * Ty - umy * Tx;
    double Tf = 2.0 / (1 + Tx * Tx + Ty * Ty + Tz * Tz);
    double Sx = Tf * Tx, Sy = Tf * Ty, Sz = Tf * Tz;
    double uxp = umx + upy * Sz - upz * Sy;
    double uyp = umy + upz * Sx - upx * Sz;
    double uzp = umz + upx * Sy - upy * Sx;
    *ux = uxp + efactor * Ex;
    *uy = uyp + efactor * Ey;
    *uz = uzp + efactor * Ez;
    *inv_gamma = 1.0 / sqrt(1 + (*ux) * (*ux) + (*uy) * (*uy) + (*uz) * (*uz));
}

static inline void shape_S(double delta, int shift, double *S) { /* calculate_S / calculate_S0, current_deposit.h:7-35 */
    double d2 = delta * delta;
    double lo = 0.5 * (d2 + delta + 0.25), mid = 0.75 - d2, hi = 0.5 * (d2 - delta + 0.25);
    S[0] = S[1] = S[2] = S[3] = S[4] = 0.0;
    S[1 + shift] = lo;
    S[2 + shift] = mid;
    S[3 + shift] = hi;
}

static inline i64 wrap_base(i64 i, i64 N) { /* current_deposit.h:417-423 */
    while (i < 0) i += N;
    while (i >= N) i -= N;
    return i;
}
static inline i64 wrap_once(i64 i, i64 N) { /* current_deposit.h:289-291 */
    if (i < 0) return i + N;
    if (i >= N) return i - N;
    return i;
}

static void deposit_one_3d(double *rho, double *jx, double *jy, double *jz, double x, double y, double z,
                           double ux, double uy, double uz, double inv_gamma, i64 NX, i64 NY, i64 NZ,
                           double dx, double dy, double dz, double x0, double y0, double z0, double dt, double w,
                           double q_dV, double q_dydzdt, double q_dxdzdt, double q_dxdydt) {
    /* current_deposit_3d_fast + _cells, current_deposit.h:275-440 */
    double vx = ux * C_LIGHT * inv_gamma, vy = uy * C_LIGHT * inv_gamma, vz = uz * C_LIGHT * inv_gamma;
    double xo = x - vx * 0.5 * dt - x0, xa = x + vx * 0.5 * dt - x0;
    double yo = y - vy * 0.5 * dt - y0, ya = y + vy * 0.5 * dt - y0;
    double zo = z - vz * 0.5 * dt - z0, za = z + vz * 0.5 * dt - z0;
    double X0 = xo / dx, Y0 = yo / dy, Z0 = zo / dz, X1 = xa / dx, Y1 = ya / dy, Z1 = za / dz;
    int ix0 = (int)floor(X0 + 0.5), iy0 = (int)floor(Y0 + 0.5), iz0 = (int)floor(Z0 + 0.5);
    int ix1 = (int)floor(X1 + 0.5), iy1 = (int)floor(Y1 + 0.5), iz1 = (int)floor(Z1 + 0.5);
    int dcx = ix1 - ix0, dcy = iy1 - iy0, dcz = iz1 - iz0;
    double S0x[5], S0y[5], S0z[5], S1x[5], S1y[5], S1z[5], DSx[5], DSy[5], DSz[5];
    shape_S(ix0 - X0, 0, S0x); shape_S(iy0 - Y0, 0, S0y); shape_S(iz0 - Z0, 0, S0z);
    shape_S(ix1 - X1, dcx, S1x); shape_S(iy1 - Y1, dcy, S1y); shape_S(iz1 - Z1, dcz, S1z);
    for (int i = 0; i < 5; i++) { DSx[i] = S1x[i] - S0x[i]; DSy[i] = S1y[i] - S0y[i]; DSz[i] = S1z[i] - S0z[i]; }
    double cd = q_dV * w, fdx = q_dydzdt * w, fdy = q_dxdzdt * w, fdz = q_dxdydt * w;
    int is = dcx < 0 ? 0 : 1, ie = dcx > 0 ? 5 : 4, js = dcy < 0 ? 0 : 1, je = dcy > 0 ? 5 : 4;
    int ks = dcz < 0 ? 0 : 1, ke = dcz > 0 ? 5 : 4;
    i64 bx0 = wrap_base(ix0, NX), by0 = wrap_base(iy0, NY), bz0 = wrap_base(iz0, NZ);
    double jxb[5][5] = {{0}};
    for (int i = is; i < ie; i++) {
        i64 ix = wrap_once(bx0 + i - 2, NX);
        double ax = S0x[i] + 0.5 * DSx[i], cx = 0.5 * S0x[i] + ONE_THIRD * DSx[i], fx = fdx * DSx[i];
        double jyb[5] = {0};
        for (int j = js; j < je; j++) {
            i64 iy = wrap_once(by0 + j - 2, NY);
            double ay = S0y[j] + 0.5 * DSy[j], cy = 0.5 * S0y[j] + ONE_THIRD * DSy[j], fy = fdy * DSy[j];
            double tz = ax * S0y[j] + cx * DSy[j];
            double jzb = 0;
            for (int k = ks; k < ke; k++) {
                i64 iz = wrap_once(bz0 + k - 2, NZ);
                double tx = ay * S0z[k] + cy * DSz[k];
                double ty = ax * S0z[k] + cx * DSz[k];
                jxb[k][j] -= fx * tx;
                jyb[k] -= fy * ty;
                jzb -= fdz * DSz[k] * tz;
                i64 id = iz + iy * NZ + ix * NY * NZ;
                jx[id] += jxb[k][j];
                jy[id] += jyb[k];
                jz[id] += jzb;
                rho[id] += cd * S1x[i] * S1y[j] * S1z[k];
            }
        }
    }
}

static void deposit_one_2d(double *rho, double *jx, double *jy, double *jz, double x, double y,
                           double ux, double uy, double uz, double inv_gamma, i64 NX, i64 NY,
                           double dx, double dy, double x0, double y0, double dt, double w,
                           double q_dxdy, double q_dydt, double q_dxdt) {
    /* current_deposit_2d_fast + _cells, current_deposit.h:150-268 */
    double vx = ux * C_LIGHT * inv_gamma, vy = uy * C_LIGHT * inv_gamma, vz = uz * C_LIGHT * inv_gamma;
    double xo = x - vx * 0.5 * dt - x0, yo = y - vy * 0.5 * dt - y0;
    double xa = x + vx * 0.5 * dt - x0, ya = y + vy * 0.5 * dt - y0;
    double X0 = xo / dx, Y0 = yo / dy, X1 = xa / dx, Y1 = ya / dy;
    int ix0 = (int)floor(X0 + 0.5), iy0 = (int)floor(Y0 + 0.5), ix1 = (int)floor(X1 + 0.5), iy1 = (int)floor(Y1 + 0.5);
    int dcx = ix1 - ix0, dcy = iy1 - iy0;
    double S0x[5], S0y[5], S1x[5], S1y[5], DSx[5], DSy[5];
    shape_S(ix0 - X0, 0, S0x); shape_S(iy0 - Y0, 0, S0y);
    shape_S(ix1 - X1, dcx, S1x); shape_S(iy1 - Y1, dcy, S1y);
    for (int i = 0; i < 5; i++) { DSx[i] = S1x[i] - S0x[i]; DSy[i] = S1y[i] - S0y[i]; }
    double cd = q_dxdy * w, fdx = q_dydt * w, fdy = q_dxdt * w, fvz = cd * vz;
    const double one_twelfth = 1.0 / 12.0;
    int is = dcx < 0 ? 0 : 1, ie = dcx > 0 ? 5 : 4, js = dcy < 0 ? 0 : 1, je = dcy > 0 ? 5 : 4;
    i64 ixs[5], iys[5]; /* PRECOMPUTE_WRAP_INDICES, current_deposit.h:41-49 */
    for (int t = 0; t < 5; t++) { ixs[t] = wrap_base((i64)ix0 + t - 2, NX); iys[t] = wrap_base((i64)iy0 + t - 2, NY); }
    double jxb[5] = {0, 0, 0, 0, 0};
    for (int i = is; i < ie; i++) {
        double jyb = 0.0;
        double a = S0x[i] + 0.5 * DSx[i], fxi = fdx * DSx[i], t12 = one_twelfth * DSx[i];
        for (int j = js; j < je; j++) {
            double b = S0y[j] + 0.5 * DSy[j];
            double wy = DSy[j] * a;
            double wz = a * b + t12 * DSy[j];
            jxb[j] -= fxi * b;
            jyb -= fdy * wy;
            i64 id = iys[j] + ixs[i] * NY;
            jx[id] += jxb[j];
            jy[id] += jyb;
            jz[id] += fvz * wz;
            rho[id] += cd * S1x[i] * S1y[j];
        }
    }
}

/* One species, one patch: half push, gather, Boris, half push, deposit.  unified_pusher_3d.c:281-431.
 * Dead or NaN-position particles are skipped (:336-339,387-388).  part[] = the six *_part arrays (written). */
void orc_push_deposit_3d(double *x, double *y, double *z, double *ux, double *uy, double *uz, double *inv_gamma,
                         const double *w, const u8 *is_dead, double **part, i64 npart,
                         const double *ex, const double *ey, const double *ez, const double *bx, const double *by,
                         const double *bz, double *jx, double *jy, double *jz, double *rho,
                         i64 nx, i64 ny, i64 nz, i64 ng, double dx, double dy, double dz,
                         double x0, double y0, double z0, double dt, double q, double m) {
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng, NZ = nz + 2 * ng;
    const double efactor = q * dt / (2 * m * C_LIGHT), bfactor = q * dt / (2 * m), cdt = C_LIGHT * 0.5 * dt;
    const double q_dV = q / (dx * dy * dz), q_dydzdt = q / (dy * dz * dt), q_dxdzdt = q / (dx * dz * dt),
                 q_dxdydt = q / (dx * dy * dt);
    const double idx = 1.0 / dx, idy = 1.0 / dy, idz = 1.0 / dz;
    for (i64 ip = 0; ip < npart; ip++) {
        if (is_dead[ip] || isnan(x[ip]) || isnan(y[ip]) || isnan(z[ip])) continue;
        x[ip] += cdt * inv_gamma[ip] * ux[ip];
        y[ip] += cdt * inv_gamma[ip] * uy[ip];
        z[ip] += cdt * inv_gamma[ip] * uz[ip];
        double X = (x[ip] - x0) * idx, Y = (y[ip] - y0) * idy, Z = (z[ip] - z0) * idz;
        i64 ix1 = (i64)floor(X + 0.5), ix2 = (i64)floor(X), iy1 = (i64)floor(Y + 0.5), iy2 = (i64)floor(Y);
        i64 iz1 = (i64)floor(Z + 0.5), iz2 = (i64)floor(Z);
        double gx[3], gy[3], gz[3], hx[3], hy[3], hz[3];
        tsc3(ix1 - X, gx); tsc3(ix2 - X + 0.5, hx);
        tsc3(iy1 - Y, gy); tsc3(iy2 - Y + 0.5, hy);
        tsc3(iz1 - Z, gz); tsc3(iz2 - Z + 0.5, hz);
        double Ex = gather27(ex, hx, gy, gz, ix2, iy1, iz1, NX, NY, NZ);
        double Ey = gather27(ey, gx, hy, gz, ix1, iy2, iz1, NX, NY, NZ);
        double Ez = gather27(ez, gx, gy, hz, ix1, iy1, iz2, NX, NY, NZ);
        double Bx = gather27(bx, gx, hy, hz, ix1, iy2, iz2, NX, NY, NZ);
        double By = gather27(by, hx, gy, hz, ix2, iy1, iz2, NX, NY, NZ);
        double Bz = gather27(bz, hx, hy, gz, ix2, iy2, iz1, NX, NY, NZ);
        if (part) { part[0][ip] = Ex; part[1][ip] = Ey; part[2][ip] = Ez; part[3][ip] = Bx; part[4][ip] = By; part[5][ip] = Bz; }
        boris_kick(&ux[ip], &uy[ip], &uz[ip], &inv_gamma[ip], Ex, Ey, Ez, Bx, By, Bz, efactor, bfactor);
        x[ip] += cdt * inv_gamma[ip] * ux[ip];
        y[ip] += cdt * inv_gamma[ip] * uy[ip];
        z[ip] += cdt * inv_gamma[ip] * uz[ip];
    }
    for (i64 ip = 0; ip < npart; ip++) {
        if (is_dead[ip] || isnan(x[ip]) || isnan(y[ip]) || isnan(z[ip])) continue;
        deposit_one_3d(rho, jx, jy, jz, x[ip], y[ip], z[ip], ux[ip], uy[ip], uz[ip], inv_gamma[ip], NX, NY, NZ,
                       dx, dy, dz, x0, y0, z0, dt, w[ip], q_dV, q_dydzdt, q_dxdzdt, q_dxdydt);
    }
}

/* 2D twin: unified_pusher_2d.c:157-365 */
void orc_push_deposit_2d(double *x, double *y, double *ux, double *uy, double *uz, double *inv_gamma,
                         const double *w, const u8 *is_dead, double **part, i64 npart,
                         const double *ex, const double *ey, const double *ez, const double *bx, const double *by,
                         const double *bz, double *jx, double *jy, double *jz, double *rho,
                         i64 nx, i64 ny, i64 ng, double dx, double dy, double x0, double y0, double dt, double q, double m) {
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng;
    const double efactor = q * dt / (2 * m * C_LIGHT), bfactor = q * dt / (2 * m), cdt = C_LIGHT * 0.5 * dt;
    const double q_dxdy = q / (dx * dy), q_dydt = q / (dy * dt), q_dxdt = q / (dx * dt);
    const double idx = 1.0 / dx, idy = 1.0 / dy;
    for (i64 ip = 0; ip < npart; ip++) {
        if (is_dead[ip] || isnan(x[ip]) || isnan(y[ip])) continue;
        x[ip] += cdt * inv_gamma[ip] * ux[ip];
        y[ip] += cdt * inv_gamma[ip] * uy[ip];
        double X = (x[ip] - x0) * idx, Y = (y[ip] - y0) * idy;
        i64 ix1 = (i64)floor(X + 0.5), ix2 = (i64)floor(X), iy1 = (i64)floor(Y + 0.5), iy2 = (i64)floor(Y);
        double gx[3], gy[3], hx[3], hy[3];
        tsc3(ix1 - X, gx); tsc3(ix2 - X + 0.5, hx);
        tsc3(iy1 - Y, gy); tsc3(iy2 - Y + 0.5, hy);
        double Ex = gather9(ex, hx, gy, ix2, iy1, NX, NY);
        double Ey = gather9(ey, gx, hy, ix1, iy2, NX, NY);
        double Ez = gather9(ez, gx, gy, ix1, iy1, NX, NY);
        double Bx = gather9(bx, gx, hy, ix1, iy2, NX, NY);
        double By = gather9(by, hx, gy, ix2, iy1, NX, NY);
        double Bz = gather9(bz, hx, hy, ix2, iy2, NX, NY);
        if (part) { part[0][ip] = Ex; part[1][ip] = Ey; part[2][ip] = Ez; part[3][ip] = Bx; part[4][ip] = By; part[5][ip] = Bz; }
        boris_kick(&ux[ip], &uy[ip], &uz[ip], &inv_gamma[ip], Ex, Ey, Ez, Bx, By, Bz, efactor, bfactor);
        x[ip] += cdt * inv_gamma[ip] * ux[ip];
        y[ip] += cdt * inv_gamma[ip] * uy[ip];
    }
    for (i64 ip = 0; ip < npart; ip++) {
        if (is_dead[ip] || isnan(x[ip]) || isnan(y[ip])) continue;
        deposit_one_2d(rho, jx, jy, jz, x[ip], y[ip], ux[ip], uy[ip], uz[ip], inv_gamma[ip], NX, NY,
                       dx, dy, x0, y0, dt, w[ip], q_dxdy, q_dydt, q_dxdt);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Per-patch bucket sort.  core/sort/cpu3d.c:8-156 (2D twin: cpu2d.c; pass nz_b=1, z=NULL).
 * Returns nbuf (number of slots rewritten).  attrs = nattrs fp64 arrays permuted together with is_dead.
 * Work arrays (all caller-provided, as in the reference's sorter object):
 *   bucket_count/bound_min/bound_max [nbin], particle_index/particle_index_ref/particle_index_target [npart]
 * ---------------------------------------------------------------------------------------------- */
i64 orc_sort_patch(const double *x, const double *y, const double *z, u8 *is_dead, double **attrs, i64 nattrs,
                   i64 npart, i64 nxb, i64 nyb, i64 nzb, double dxb, double dyb, double dzb,
                   double x0, double y0, double z0, int reverse_x,
                   i64 *bucket_count, i64 *bound_min, i64 *bound_max,
                   i64 *pidx, i64 *pref, i64 *ptarget) {
    const i64 nbin = nxb * nyb * nzb;
    memset(bucket_count, 0, sizeof(i64) * nbin);
    i64 icell = 0;
    for (i64 ip = 0; ip < npart; ip++) { /* calculate_cell_index, cpu3d.c:8-56 */
        if (!is_dead[ip]) {
            i64 ix = (i64)floor((x[ip] - x0) / dxb);
            i64 iy = (i64)floor((y[ip] - y0) / dyb);
            i64 iz = z ? (i64)floor((z[ip] - z0) / dzb) : 0;
            if (reverse_x) {
                if (ix < 0) ix = 0; else if (ix >= nxb) ix = nxb - 1;
                if (iy < 0) iy = 0; else if (iy >= nyb) iy = nyb - 1;
                if (iz < 0) iz = 0; else if (iz >= nzb) iz = nzb - 1;
                icell = iz + iy * nzb + (nxb - 1 - ix) * nyb * nzb;
            } else {
                icell = iz + iy * nzb + ix * nyb * nzb;
                if (!(0 <= ix && ix < nxb && 0 <= iy && iy < nyb && 0 <= iz && iz < nzb)) icell = nbin - 1;
            }
        } /* dead slots inherit the bucket of the previous slot (cpu3d.c:46-54) */
        pidx[ip] = icell;
        bucket_count[icell] += 1;
    }
    i64 run = 0;
    for (i64 b = 0; b < nbin; b++) { /* calculate_bucket_bound, cpu3d.c:58-71 */
        bound_min[b] = run;
        run += bucket_count[b];
        bound_max[b] = run;
    }
    for (i64 b = 0; b < nbin; b++) /* cpu3d.c:83-94 */
        for (i64 ip = bound_min[b]; ip < bound_max[b]; ip++) pref[ip] = b;
    i64 nbuf = 0;
    i64 *cnt_not = (i64 *)calloc(nbin, sizeof(i64));
    for (i64 ip = 0; ip < npart; ip++)
        if (pidx[ip] != pref[ip]) { cnt_not[pref[ip]]++; ptarget[nbuf++] = ip; }
    if (nbuf == 0) { free(cnt_not); return 0; }
    i64 *start = (i64 *)malloc(sizeof(i64) * nbin);
    i64 *dest = (i64 *)malloc(sizeof(i64) * nbuf);
    /* position of each misplaced slot's value in the staging buffer: stable by owning bucket (cpu3d.c:111-128) */
    start[0] = 0;
    for (i64 b = 1; b < nbin; b++) start[b] = start[b - 1] + cnt_not[b - 1];
    for (i64 i = 0; i < nbuf; i++) dest[i] = start[pidx[ptarget[i]]]++;
    double *buf = (double *)malloc(sizeof(double) * nbuf);
    for (i64 a = 0; a < nattrs; a++) {
        for (i64 i = 0; i < nbuf; i++) buf[dest[i]] = attrs[a][ptarget[i]];
        for (i64 i = 0; i < nbuf; i++) attrs[a][ptarget[i]] = buf[i];
    }
    u8 *dbuf = (u8 *)buf;
    for (i64 i = 0; i < nbuf; i++) dbuf[dest[i]] = is_dead[ptarget[i]];
    for (i64 i = 0; i < nbuf; i++) is_dead[ptarget[i]] = dbuf[i];
    free(buf); free(dest); free(start); free(cnt_not);
    return nbuf;
}

/* ------------------------------------------------------------------------------------------------
 * Intra-rank particle migration.  core/patch/sync_particles_3d.c:79-193 (classify), :365-482 (counts and the
 * growth rule), :484-695 (fill), 2D twin sync_particles_2d.c.
 * box[p*6 + {0..5}] = xmin,xmax,ymin,ymax,zmin,zmax already widened by half a cell (:402-409).
 * ---------------------------------------------------------------------------------------------- */
static int classify(double x, double y, double z, const double *bx, int dim) { /* sync_particles_3d.c:85-192 */
    int sx = x < bx[0] ? -1 : (x > bx[1] ? 1 : 0);
    int sy = y < bx[2] ? -1 : (y > bx[3] ? 1 : 0);
    int sz = dim == 3 ? (z < bx[4] ? -1 : (z > bx[5] ? 1 : 0)) : 0;
    if (!sx && !sy && !sz) return -1;
    return dim == 3 ? find_dir(DIR3, 26, sx, sy, sz) : find_dir(DIR2, 8, sx, sy, sz);
}

/* out: npart_out[npatch*nb], npart_incoming[npatch], npart_to_extend[npatch], npart_alive[npatch] */
void orc_migrate_count(double **x, double **y, double **z, u8 **is_dead, const i64 *npart, i64 npatch, i64 dim,
                       const double *box, const i64 *nbr, i64 *npart_out, i64 *npart_incoming, i64 *npart_to_extend,
                       i64 *npart_alive) {
    const int nb = dim == 3 ? 26 : 8;
    const int (*tab)[3] = dim == 3 ? DIR3 : DIR2;
    memset(npart_out, 0, sizeof(i64) * npatch * nb);
    for (i64 p = 0; p < npatch; p++)
        for (i64 ip = 0; ip < npart[p]; ip++) {
            if (is_dead[p][ip]) continue;
            int b = classify(x[p][ip], y[p][ip], dim == 3 ? z[p][ip] : 0.0, box + 6 * p, (int)dim);
            if (b >= 0) npart_out[p * nb + b]++;
        }
    for (i64 p = 0; p < npatch; p++) {
        i64 incoming = 0, ndead = 0;
        for (int b = 0; b < nb; b++) {
            i64 q = nbr[p * nb + b];
            if (q >= 0) incoming += npart_out[q * nb + opposite(tab, nb, b)];
        }
        for (i64 ip = 0; ip < npart[p]; ip++) ndead += is_dead[p][ip] ? 1 : 0;
        npart_alive[p] = npart[p] - ndead + incoming;
        npart_to_extend[p] = 0;
        if (incoming - ndead > 0) npart_to_extend[p] = incoming - ndead + (i64)(npart[p] * 0.25); /* :468-473 */
        npart_incoming[p] = incoming;
    }
}

/* attrs[p*nattrs + a]; ia_x/ia_y/ia_z = positions of x,y,z inside the attr list (ia_z < 0 in 2D).
 * glob = xmin,xmax,ymin,ymax,zmin,zmax of the global box.  nan_dead_positions: 3D reference NaNs the position
 * of every dead slot (:333-338); the 2D reference does not (sync_particles_2d.c:185-202). */
void orc_migrate_fill(double **attrs, i64 nattrs, i64 ia_x, i64 ia_y, i64 ia_z, u8 **is_dead, const i64 *npart,
                      i64 npatch, i64 dim, const double *box, const i64 *nbr, const double *glob,
                      double dx, double dy, double dz, const i64 *npart_incoming) {
    const int nb = dim == 3 ? 26 : 8;
    const int (*tab)[3] = dim == 3 ? DIR3 : DIR2;
    /* leaver lists per (source patch, direction), ascending slot order (get_incoming_index, :204-299) */
    i64 **lists = (i64 **)calloc(npatch * nb, sizeof(i64 *));
    i64 *cnt = (i64 *)calloc(npatch * nb, sizeof(i64));
    for (i64 p = 0; p < npatch; p++) {
        double *x = attrs[p * nattrs + ia_x], *y = attrs[p * nattrs + ia_y], *z = ia_z >= 0 ? attrs[p * nattrs + ia_z] : NULL;
        for (int b = 0; b < nb; b++) lists[p * nb + b] = (i64 *)malloc(sizeof(i64) * (npart[p] + 1));
        for (i64 ip = 0; ip < npart[p]; ip++) {
            if (is_dead[p][ip]) continue;
            int b = classify(x[ip], y[ip], dim == 3 ? z[ip] : 0.0, box + 6 * p, (int)dim);
            if (b >= 0 && nbr[p * nb + b] >= 0) lists[p * nb + b][cnt[p * nb + b]++] = ip;
        }
    }
    const double L[3] = {glob[1] - glob[0], glob[3] - glob[2], glob[5] - glob[4]};
    const double cell[3] = {dx, dy, dz};
    const i64 ia[3] = {ia_x, ia_y, dim == 3 ? ia_z : -1};
    for (i64 p = 0; p < npatch; p++) {
        i64 nnew = npart_incoming[p];
        if (nnew <= 0) continue;
        double *buf = (double *)malloc(sizeof(double) * nattrs * nnew);
        i64 nb_filled = 0;
        for (int b = 0; b < nb; b++) { /* fill_boundary_particles_to_buffer, :302-323 */
            i64 q = nbr[p * nb + b];
            if (q < 0) continue;
            int ob = opposite(tab, nb, b);
            for (i64 t = 0; t < cnt[q * nb + ob]; t++, nb_filled++)
                for (i64 a = 0; a < nattrs; a++) buf[nb_filled * nattrs + a] = attrs[q * nattrs + a][lists[q * nb + ob][t]];
        }
        i64 ib = 0;
        for (i64 ip = 0; ip < npart[p] && ib < nnew; ip++) { /* :635-671 */
            if (!is_dead[p][ip]) continue;
            for (int d = 0; d < 3; d++) { /* handle_periodic, :349-363 */
                if (ia[d] < 0) continue;
                double *c = &buf[ib * nattrs + ia[d]], c0 = *c;
                if (c0 > glob[2 * d + 1] && fabs(box[6 * p + 2 * d] - glob[2 * d]) < cell[d]) *c -= L[d];
                if (c0 < glob[2 * d] && fabs(box[6 * p + 2 * d + 1] - glob[2 * d + 1]) < cell[d]) *c += L[d];
            }
            for (i64 a = 0; a < nattrs; a++) attrs[p * nattrs + a][ip] = buf[ib * nattrs + a];
            is_dead[p][ip] = 0;
            ib++;
        }
        free(buf);
    }
    for (i64 p = 0; p < npatch; p++) { /* mark_out_of_bound_as_dead, :326-346 */
        double *x = attrs[p * nattrs + ia_x], *y = attrs[p * nattrs + ia_y], *z = ia_z >= 0 ? attrs[p * nattrs + ia_z] : NULL;
        const double *bx = box + 6 * p;
        for (i64 ip = 0; ip < npart[p]; ip++) {
            int out = 0;
            if (is_dead[p][ip]) { if (dim != 3) continue; out = 1; }
            else if (classify(x[ip], y[ip], dim == 3 ? z[ip] : 0.0, bx, (int)dim) >= 0) { out = 1; is_dead[p][ip] = 1; }
            if (out) { x[ip] = NAN; y[ip] = NAN; if (dim == 3) z[ip] = NAN; }
        }
    }
    for (i64 i = 0; i < npatch * nb; i++) free(lists[i]);
    free(lists); free(cnt);
}

/* ------------------------------------------------------------------------------------------------
 * CPML (convolutional PML) edge patches: kappa-scaled Yee update and the psi auxiliary currents.
 * core/boundary/cpml.py:342-362 (2D), :437-477 (3D), :527-730 (psi); coefficients :118-126.
 * Expression order is the reference's (numba, no contraction).  kappa_* are per-patch 1-D arrays along
 * each axis (1.0 outside the layer).  2D and 3D associate the products differently, as the reference does.
 * ---------------------------------------------------------------------------------------------- */
void orc_update_efield_cpml_3d(double *ex, double *ey, double *ez, const double *bx, const double *by, const double *bz,
                               const double *jx, const double *jy, const double *jz, const double *kex, const double *key,
                               const double *kez, i64 nx, i64 ny, i64 nz, i64 ng, double dx, double dy, double dz,
                               double bfactor, double jfactor) {
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng, NZ = nz + 2 * ng;
    for (i64 i = 0; i < nx; i++) {
        const double bfx = bfactor / kex[i];
        for (i64 j = 0; j < ny; j++) {
            const double bfy = bfactor / key[j];
            for (i64 k = 0; k < nz; k++) {
                const double bfz = bfactor / kez[k];
                i64 c = IX3(i, j, k), xm = IX3(i - 1, j, k), ym = IX3(i, j - 1, k), zm = IX3(i, j, k - 1);
                ex[c] += (bfy * (bz[c] - bz[ym]) / dy - bfz * (by[c] - by[zm]) / dz) - jfactor * jx[c];
                ey[c] += (bfz * (bx[c] - bx[zm]) / dz - bfx * (bz[c] - bz[xm]) / dx) - jfactor * jy[c];
                ez[c] += (bfx * (by[c] - by[xm]) / dx - bfy * (bx[c] - bx[ym]) / dy) - jfactor * jz[c];
            }
        }
    }
}

void orc_update_bfield_cpml_3d(const double *ex, const double *ey, const double *ez, double *bx, double *by, double *bz,
                               const double *kbx, const double *kby, const double *kbz, i64 nx, i64 ny, i64 nz, i64 ng,
                               double dx, double dy, double dz, double dt) {
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng, NZ = nz + 2 * ng;
    for (i64 i = 0; i < nx; i++) {
        const double efx = dt / kbx[i];
        for (i64 j = 0; j < ny; j++) {
            const double efy = dt / kby[j];
            for (i64 k = 0; k < nz; k++) {
                const double efz = dt / kbz[k];
                i64 c = IX3(i, j, k), xp = IX3(i + 1, j, k), yp = IX3(i, j + 1, k), zp = IX3(i, j, k + 1);
                bx[c] -= (efy * (ez[yp] - ez[c]) / dy - efz * (ey[zp] - ey[c]) / dz);
                by[c] -= (efz * (ex[zp] - ex[c]) / dz - efx * (ez[xp] - ez[c]) / dx);
                bz[c] -= (efx * (ey[xp] - ey[c]) / dx - efy * (ex[yp] - ex[c]) / dy);
            }
        }
    }
}

void orc_update_efield_cpml_2d(double *ex, double *ey, double *ez, const double *bx, const double *by, const double *bz,
                               const double *jx, const double *jy, const double *jz, const double *kex, const double *key,
                               i64 nx, i64 ny, i64 ng, double dx, double dy, double bfactor, double jfactor) {
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng;
    for (i64 i = 0; i < nx; i++) {
        const double bfx = bfactor / kex[i];
        for (i64 j = 0; j < ny; j++) {
            const double bfy = bfactor / key[j];
            i64 c = IX2(i, j), xm = IX2(i - 1, j), ym = IX2(i, j - 1);
            ex[c] += bfy * ((bz[c] - bz[ym]) / dy) - jfactor * jx[c];
            ey[c] += bfx * (-(bz[c] - bz[xm]) / dx) - jfactor * jy[c];
            ez[c] += bfx * ((by[c] - by[xm]) / dx) - bfy * ((bx[c] - bx[ym]) / dy) - jfactor * jz[c];
        }
    }
}

void orc_update_bfield_cpml_2d(const double *ex, const double *ey, const double *ez, double *bx, double *by, double *bz,
                               const double *kbx, const double *kby, i64 nx, i64 ny, i64 ng, double dx, double dy, double dt) {
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng;
    for (i64 i = 0; i < nx; i++) {
        const double efx = dt / kbx[i];
        for (i64 j = 0; j < ny; j++) {
            const double efy = dt / kby[j];
            i64 c = IX2(i, j), xp = IX2(i + 1, j), yp = IX2(i, j + 1);
            bx[c] -= efy * ((ez[yp] - ez[c]) / dy);
            by[c] -= efx * (-(ez[xp] - ez[c]) / dx);
            bz[c] -= efx * ((ey[xp] - ey[c]) / dx) - efy * ((ex[yp] - ex[c]) / dy);
        }
    }
}

/* psi update of one PML face and the correction of the two field components it drives.
 * axis 0/1/2 = x/y/z.  is_b = 0: E-side (backward difference of B, fac = dt c^2); 1: B-side (forward difference of E,
 * fac = dt).  f1/f2: the two corrected components, s1/s2 their signs, g1/g2: the differenced source components:
 *   E: x: (ey,-,bz) (ez,+,by)   y: (ex,+,bz) (ez,-,bx)   z: (ex,-,by) (ey,+,bx)
 *   B: x: (by,+,ez) (bz,-,ey)   y: (bx,-,ez) (bz,+,ex)   z: (bx,+,ey) (by,-,ex)
 * psi arrays have the interior shape (nx, ny[, nz]).  kappa/sigma/a: 1-D along the axis. */
void orc_update_psi(double *f1, double *f2, const double *g1, const double *g2, double *psi1, double *psi2, double s1, double s2,
                    const double *kappa, const double *sigma, const double *a, i64 axis, i64 is_b, i64 start, i64 stop,
                    i64 dim, i64 nx, i64 ny, i64 nz, i64 ng, double d, double dt) {
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng, NZ = dim == 3 ? nz + 2 * ng : 1;
    if (dim == 2) nz = 1;
    const double fac = is_b ? dt : dt * C_LIGHT * C_LIGHT;
    const i64 n[3] = {nx, ny, nz};
    for (i64 i = 0; i < nx; i++)
        for (i64 j = 0; j < ny; j++)
            for (i64 k = 0; k < nz; k++) {
                const i64 idx[3] = {i, j, k};
                const i64 ipos = idx[axis];
                if (ipos < start || ipos >= stop) continue;
                const double kap = kappa[ipos], sig = sigma[ipos], ac = a[ipos];
                const double bcoeff = exp(-(sig / kap + ac) * dt);
                const double ccoeff = (bcoeff - 1) * sig / kap / (sig + kap * ac) / d;
                i64 o[3] = {i, j, k};
                o[axis] += is_b ? 1 : -1;
                const i64 c = wrapneg(k, NZ) + wrapneg(j, NY) * NZ + wrapneg(i, NX) * NY * NZ;
                const i64 nb = wrapneg(o[2], NZ) + wrapneg(o[1], NY) * NZ + wrapneg(o[0], NX) * NY * NZ;
                const i64 pi = k + n[2] * (j + n[1] * i);
                const double d1 = is_b ? g1[nb] - g1[c] : g1[c] - g1[nb];
                const double d2 = is_b ? g2[nb] - g2[c] : g2[c] - g2[nb];
                psi1[pi] = bcoeff * psi1[pi] + ccoeff * d1;
                psi2[pi] = bcoeff * psi2[pi] + ccoeff * d2;
                f1[c] += s1 * (fac * psi1[pi]);
                f2[c] += s2 * (fac * psi2[pi]);
            }
}

/* ------------------------------------------------------------------------------------------------
 * Laser antenna at xmin: B at the plane laserpos-1 of an xmin edge patch is rewritten from the source fields.
 * callback/laser.py:17-45 (2D), :47-77 (3D).  Expression order as the reference (numba, no contraction).
 * ey_src / ez_src have the shape of one x-plane of the padded grid: (NY) or (NY, NZ), wrapped indexing.
 * ---------------------------------------------------------------------------------------------- */
void orc_laser_bfields(const double *ex, const double *ey, const double *ez, double *bx, double *by, double *bz,
                       const double *jx, const double *jy, const double *jz, i64 dim, i64 nx, i64 ny, i64 nz, i64 ng,
                       double dx, double dy, double dz, double dt, i64 laserpos, i64 iy0, i64 iy1, i64 iz0, i64 iz1,
                       const double *ey_src, const double *ez_src) {
    const double c = C_LIGHT, eps0 = 8.8541878188e-12;
    const i64 NX = nx + 2 * ng, NY = ny + 2 * ng, NZ = dim == 3 ? nz + 2 * ng : 1;
    if (dim == 2) { iz0 = 0; iz1 = 1; }
    for (i64 iy = iy0; iy < iy1; iy++) /* bx[laserpos-1, iy, :] = bx[0, iy, :] (the whole padded z row in 3D) */
        for (i64 sk = 0; sk < NZ; sk++)
            bx[sk + NZ * (wrapneg(iy, NY) + NY * wrapneg(laserpos - 1, NX))] = bx[sk + NZ * (wrapneg(iy, NY) + NY * 0)];
    const double inv = 1 / ((c * dt / dx + 1) * c);
    for (i64 iy = iy0; iy < iy1; iy++)
        for (i64 iz = iz0; iz < iz1; iz++) {
#define AT(i, j, k) (wrapneg(k, NZ) + NZ * (wrapneg(j, NY) + NY * wrapneg(i, NX)))
            const i64 s = wrapneg(iz, NZ) + NZ * wrapneg(iy, NY);
            const i64 o0 = AT(0, iy, iz), om = AT(-1, iy, iz), ol = AT(laserpos, iy, iz), ot = AT(laserpos - 1, iy, iz);
            double vz = 4 * ey_src[s] + 2 * (ey[o0] + c * 0.5 * (bz[o0] + bz[om])) - 2 * ey[ol];
            if (dim == 3) vz = vz - (dt * (c * c)) * (bx[ol] - bx[AT(laserpos, iy, iz - 1)]) / dz;
            vz = vz + dt / eps0 * jy[ol] + (c * dt / dx - 1) * c * bz[ol];
            double vy = -4 * ez_src[s] - 2 * (ez[o0] - c * 0.5 * (by[o0] + by[om])) + 2 * ez[ol] -
                        (dt * (c * c)) * (bx[ol] - bx[AT(laserpos, iy - 1, iz)]) / dy - dt / eps0 * jz[ol] +
                        (c * dt / dx - 1) * c * by[ol];
            bz[ot] = inv * vz;
            by[ot] = inv * vy;
#undef AT
        }
}
