// Device-side particle math shared by the particle kernels: patch views, quadratic-spline gather on the staggered
// Yee grid, Boris rotation, Esirkepov shape factors and the per-particle deposit.
// Reference behaviour restated (not copied): core/pusher/unified/unified_pusher_3d.c:15-217, unified_pusher_2d.c:64-155,
// core/current/current_deposit.h:7-35,150-268,275-440.
#pragma once
#include "lpic_common.cuh"

struct PatchView {
    const double *ex, *ey, *ez, *bx, *by, *bz;
    double *jx, *jy, *jz, *rho;
    double x0, y0, z0;
};

__device__ __forceinline__ PatchView patch_view(const Geom &g, double *F, const double *x0, const double *y0,
                                                const double *z0, int p) {
    const size_t stride = (size_t)g.npatch * g.ncell;
    double *base = F + (size_t)p * g.ncell;
    PatchView v;
    v.ex = base + LPIC_EX * stride; v.ey = base + LPIC_EY * stride; v.ez = base + LPIC_EZ * stride;
    v.bx = base + LPIC_BX * stride; v.by = base + LPIC_BY * stride; v.bz = base + LPIC_BZ * stride;
    v.jx = base + LPIC_JX * stride; v.jy = base + LPIC_JY * stride; v.jz = base + LPIC_JZ * stride;
    v.rho = base + LPIC_RHO * stride;
    v.x0 = x0[p]; v.y0 = y0[p]; v.z0 = z0[p];
    return v;
}

struct Slots {  // attribute arenas of one species; x..ig are indexed [slot * ps] (ps = 8: 64-byte records, 1: separate arrays)
    double *x, *y, *z, *w, *ux, *uy, *uz, *ig;
    int ps;
    double *part[6];
    const u8 *dead;
    const i64 *off, *npart;
};

__device__ __forceinline__ void tsc3(double d, double *g) {  // get_gx, unified_pusher_3d.c:65-70
    const double d2 = d * d;
    g[0] = 0.5 * (0.25 + d2 + d);
    g[1] = 0.75 - d2;
    g[2] = 0.5 * (0.25 + d2 - d);
}

// 27-point weighted sum, nesting z(y(x)) as interp_field_safe_3d (unified_pusher_3d.c:79-106).
// ox/oy/oz: storage offsets of the three stencil points along each axis (already wrapped and scaled).
__device__ __forceinline__ double gather27(const double *__restrict__ F, const double *fx, const double *fy,
                                           const double *fz, const int *ox, const int *oy, const int *oz) {
    double az[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double ay[3];
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const int base = oz[c] + oy[b];
            ay[b] = fx[0] * __ldg(F + base + ox[0]) + fx[1] * __ldg(F + base + ox[1]) + fx[2] * __ldg(F + base + ox[2]);
        }
        az[c] = fy[0] * ay[0] + fy[1] * ay[1] + fy[2] * ay[2];
    }
    return fz[0] * az[0] + fz[1] * az[1] + fz[2] * az[2];
}

__device__ __forceinline__ double gather9(const double *__restrict__ F, const double *fx, const double *fy,
                                          const int *ox, const int *oy) {
    double a[3];
#pragma unroll
    for (int b = 0; b < 3; b++)
        a[b] = fx[0] * __ldg(F + oy[b] + ox[0]) + fx[1] * __ldg(F + oy[b] + ox[1]) + fx[2] * __ldg(F + oy[b] + ox[2]);
    return fy[0] * a[0] + fy[1] * a[1] + fy[2] * a[2];
}

__device__ __forceinline__ void offsets3(int i, int N, int scale, int *o) {
    o[0] = wrapneg(i - 1, N) * scale;
    o[1] = wrapneg(i, N) * scale;
    o[2] = wrapneg(i + 1, N) * scale;
}

// E and B at the particle: ex(h,g,g) ey(g,h,g) ez(g,g,h) bx(g,h,h) by(h,g,h) bz(h,h,g)  (unified_pusher_3d.c:190-195)
template <int DIM>
__device__ __forceinline__ void gather_eb(const Geom &g, const PatchView &v, double x, double y, double z, double *eb) {
    const double X = (x - v.x0) * (1.0 / g.dx), Y = (y - v.y0) * (1.0 / g.dy);
    const double fX = floor(X), fY = floor(Y), rX = floor(X + 0.5), rY = floor(Y + 0.5);
    double gx[3], gy[3], hx[3], hy[3];
    tsc3(rX - X, gx); tsc3(fX - X + 0.5, hx);
    tsc3(rY - Y, gy); tsc3(fY - Y + 0.5, hy);
    int ogx[3], ohx[3], ogy[3], ohy[3];
    offsets3((int)rX, g.NX, g.NY * g.NZ, ogx); offsets3((int)fX, g.NX, g.NY * g.NZ, ohx);
    offsets3((int)rY, g.NY, g.NZ, ogy); offsets3((int)fY, g.NY, g.NZ, ohy);
    if (DIM == 3) {
        const double Z = (z - v.z0) * (1.0 / g.dz);
        const double fZ = floor(Z), rZ = floor(Z + 0.5);
        double gz[3], hz[3];
        tsc3(rZ - Z, gz); tsc3(fZ - Z + 0.5, hz);
        int ogz[3], ohz[3];
        offsets3((int)rZ, g.NZ, 1, ogz); offsets3((int)fZ, g.NZ, 1, ohz);
        eb[0] = gather27(v.ex, hx, gy, gz, ohx, ogy, ogz);
        eb[1] = gather27(v.ey, gx, hy, gz, ogx, ohy, ogz);
        eb[2] = gather27(v.ez, gx, gy, hz, ogx, ogy, ohz);
        eb[3] = gather27(v.bx, gx, hy, hz, ogx, ohy, ohz);
        eb[4] = gather27(v.by, hx, gy, hz, ohx, ogy, ohz);
        eb[5] = gather27(v.bz, hx, hy, gz, ohx, ohy, ogz);
    } else {
        eb[0] = gather9(v.ex, hx, gy, ohx, ogy);
        eb[1] = gather9(v.ey, gx, hy, ogx, ohy);
        eb[2] = gather9(v.ez, gx, gy, ogx, ogy);
        eb[3] = gather9(v.bx, gx, hy, ogx, ohy);
        eb[4] = gather9(v.by, hx, gy, ohx, ogy);
        eb[5] = gather9(v.bz, hx, hy, ohx, ohy);
    }
}

// unified_pusher_3d.c:15-51
__device__ __forceinline__ void boris_kick(double &ux, double &uy, double &uz, double &ig, const double *eb,
                                           double efactor, double bfactor) {
    const double umx = ux + efactor * eb[0], umy = uy + efactor * eb[1], umz = uz + efactor * eb[2];
    const double igm = 1.0 / sqrt(1.0 + umx * umx + umy * umy + umz * umz);
    const double Tx = bfactor * eb[3] * igm, Ty = bfactor * eb[4] * igm, Tz = bfactor * eb[5] * igm;
    const double upx = umx + umy * Tz - umz * Ty;
    const double upy = umy + umz * Tx - umx * Tz;
    const double upz = umz + umx * Ty - umy * Tx;
    const double Tf = 2.0 / (1.0 + Tx * Tx + Ty * Ty + Tz * Tz);
    const double Sx = Tf * Tx, Sy = Tf * Ty, Sz = Tf * Tz;
    ux = umx + upy * Sz - upz * Sy + efactor * eb[0];
    uy = umy + upz * Sx - upx * Sz + efactor * eb[1];
    uz = umz + upx * Sy - upy * Sx + efactor * eb[2];
    ig = 1.0 / sqrt(1.0 + ux * ux + uy * uy + uz * uz);
}

// calculate_S0 / calculate_S (current_deposit.h:7-35): 5-point arrays, `shift` in {-1,0,1} moves the 3 non-zeros.
__device__ __forceinline__ void shape5(double delta, int shift, double *S) {
    const double d2 = delta * delta;
    const double lo = 0.5 * (d2 + delta + 0.25), mid = 0.75 - d2, hi = 0.5 * (d2 - delta + 0.25);
    S[0] = shift < 0 ? lo : 0.0;
    S[1] = shift < 0 ? mid : (shift == 0 ? lo : 0.0);
    S[2] = shift < 0 ? hi : (shift == 0 ? mid : lo);
    S[3] = shift < 0 ? 0.0 : (shift == 0 ? hi : mid);
    S[4] = shift > 0 ? hi : 0.0;
}

__device__ __forceinline__ int wrap_base(int i, int N) {  // current_deposit.h:417-423
    i %= N;
    return i < 0 ? i + N : i;
}
__device__ __forceinline__ int wrap_once(int i, int N) { return i < 0 ? i + N : (i >= N ? i - N : i); }

struct DepositCoef3 {
    double q_dV, q_dydzdt, q_dxdzdt, q_dxdydt, dt;
};

// current_deposit_3d_fast + _cells (current_deposit.h:275-440).  x,y,z are the END-of-step positions; the
// deposit reconstructs +-dt/2 around them.  Loops are fully unrolled so every array index is static.
__device__ __forceinline__ void deposit3(const Geom &g, const PatchView &v, const DepositCoef3 &k, double x, double y,
                                         double z, double ux, double uy, double uz, double ig, double w) {
    const double vx = ux * LPIC_C_LIGHT * ig, vy = uy * LPIC_C_LIGHT * ig, vz = uz * LPIC_C_LIGHT * ig;
    const double X0 = (x - vx * 0.5 * k.dt - v.x0) / g.dx, X1 = (x + vx * 0.5 * k.dt - v.x0) / g.dx;
    const double Y0 = (y - vy * 0.5 * k.dt - v.y0) / g.dy, Y1 = (y + vy * 0.5 * k.dt - v.y0) / g.dy;
    const double Z0 = (z - vz * 0.5 * k.dt - v.z0) / g.dz, Z1 = (z + vz * 0.5 * k.dt - v.z0) / g.dz;
    const int ix0 = (int)floor(X0 + 0.5), iy0 = (int)floor(Y0 + 0.5), iz0 = (int)floor(Z0 + 0.5);
    const int ix1 = (int)floor(X1 + 0.5), iy1 = (int)floor(Y1 + 0.5), iz1 = (int)floor(Z1 + 0.5);
    const int dcx = ix1 - ix0, dcy = iy1 - iy0, dcz = iz1 - iz0;
    double S0x[5], S0y[5], S0z[5], S1x[5], S1y[5], S1z[5], DSx[5], DSy[5], DSz[5];
    shape5(ix0 - X0, 0, S0x); shape5(iy0 - Y0, 0, S0y); shape5(iz0 - Z0, 0, S0z);
    shape5(ix1 - X1, dcx, S1x); shape5(iy1 - Y1, dcy, S1y); shape5(iz1 - Z1, dcz, S1z);
#pragma unroll
    for (int i = 0; i < 5; i++) { DSx[i] = S1x[i] - S0x[i]; DSy[i] = S1y[i] - S0y[i]; DSz[i] = S1z[i] - S0z[i]; }
    const double cd = k.q_dV * w, fdx = k.q_dydzdt * w, fdy = k.q_dxdzdt * w, fdz = k.q_dxdydt * w;
    const int is = dcx < 0 ? 0 : 1, ie = dcx > 0 ? 5 : 4, js = dcy < 0 ? 0 : 1, je = dcy > 0 ? 5 : 4;
    const int ks = dcz < 0 ? 0 : 1, ke = dcz > 0 ? 5 : 4;
    const int bx0 = wrap_base(ix0, g.NX), by0 = wrap_base(iy0, g.NY), bz0 = wrap_base(iz0, g.NZ);
    double jxb[5][5];
#pragma unroll
    for (int a = 0; a < 5; a++)
#pragma unroll
        for (int b = 0; b < 5; b++) jxb[a][b] = 0.0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        if (i < is || i >= ie) continue;
        const int ox = wrap_once(bx0 + i - 2, g.NX) * g.NY * g.NZ;
        const double ax = S0x[i] + 0.5 * DSx[i], cx = 0.5 * S0x[i] + LPIC_ONE_THIRD * DSx[i], fx = fdx * DSx[i];
        double jyb[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 5; j++) {
            if (j < js || j >= je) continue;
            const int oy = ox + wrap_once(by0 + j - 2, g.NY) * g.NZ;
            const double ay = S0y[j] + 0.5 * DSy[j], cy = 0.5 * S0y[j] + LPIC_ONE_THIRD * DSy[j], fy = fdy * DSy[j];
            const double tz = ax * S0y[j] + cx * DSy[j];
            const double rxy = cd * S1x[i] * S1y[j];
            double jzb = 0.0;
#pragma unroll
            for (int kk = 0; kk < 5; kk++) {
                if (kk < ks || kk >= ke) continue;
                const int id = oy + wrap_once(bz0 + kk - 2, g.NZ);
                jxb[kk][j] -= fx * (ay * S0z[kk] + cy * DSz[kk]);
                jyb[kk] -= fy * (ax * S0z[kk] + cx * DSz[kk]);
                jzb -= fdz * DSz[kk] * tz;
                // The running sums end on sum(DS) = 0: the last jx plane / jy row / jz column of the support hold only the
                // rounding residue of that cancellation (<= 4 eps of the particle's largest term) and are not deposited, like
                // exact zeros (rho and J vanish where only S0 or only S1 has support).  This kernel is bound by the L2
                // atomic unit (ncu: lts__d_atomic_input_cycles_active 47 %), so every RED that is not issued counts.
                const double rho_v = rxy * S1z[kk];
                if (i != ie - 1 && jxb[kk][j] != 0.0) atomicAdd(v.jx + id, jxb[kk][j]);
                if (j != je - 1 && jyb[kk] != 0.0) atomicAdd(v.jy + id, jyb[kk]);
                if (kk != ke - 1 && jzb != 0.0) atomicAdd(v.jz + id, jzb);
                if (rho_v != 0.0) atomicAdd(v.rho + id, rho_v);
            }
        }
    }
}

struct DepositCoef2 {
    double q_dxdy, q_dydt, q_dxdt, dt;
};

// current_deposit_2d_fast + _cells (current_deposit.h:150-268)
__device__ __forceinline__ void deposit2(const Geom &g, const PatchView &v, const DepositCoef2 &k, double x, double y,
                                         double ux, double uy, double uz, double ig, double w) {
    const double vx = ux * LPIC_C_LIGHT * ig, vy = uy * LPIC_C_LIGHT * ig, vz = uz * LPIC_C_LIGHT * ig;
    const double X0 = (x - vx * 0.5 * k.dt - v.x0) / g.dx, X1 = (x + vx * 0.5 * k.dt - v.x0) / g.dx;
    const double Y0 = (y - vy * 0.5 * k.dt - v.y0) / g.dy, Y1 = (y + vy * 0.5 * k.dt - v.y0) / g.dy;
    const int ix0 = (int)floor(X0 + 0.5), iy0 = (int)floor(Y0 + 0.5);
    const int ix1 = (int)floor(X1 + 0.5), iy1 = (int)floor(Y1 + 0.5);
    const int dcx = ix1 - ix0, dcy = iy1 - iy0;
    double S0x[5], S0y[5], S1x[5], S1y[5], DSx[5], DSy[5];
    shape5(ix0 - X0, 0, S0x); shape5(iy0 - Y0, 0, S0y);
    shape5(ix1 - X1, dcx, S1x); shape5(iy1 - Y1, dcy, S1y);
#pragma unroll
    for (int i = 0; i < 5; i++) { DSx[i] = S1x[i] - S0x[i]; DSy[i] = S1y[i] - S0y[i]; }
    const double cd = k.q_dxdy * w, fdx = k.q_dydt * w, fdy = k.q_dxdt * w, fvz = cd * vz;
    const double one_twelfth = 1.0 / 12.0;
    const int is = dcx < 0 ? 0 : 1, ie = dcx > 0 ? 5 : 4, js = dcy < 0 ? 0 : 1, je = dcy > 0 ? 5 : 4;
    double jxb[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 5; i++) {
        if (i < is || i >= ie) continue;
        const int ox = wrap_base(ix0 + i - 2, g.NX) * g.NY;
        const double a = S0x[i] + 0.5 * DSx[i], fxi = fdx * DSx[i], t12 = one_twelfth * DSx[i];
        double jyb = 0.0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            if (j < js || j >= je) continue;
            const int id = ox + wrap_base(iy0 + j - 2, g.NY);
            const double b = S0y[j] + 0.5 * DSy[j];
            jxb[j] -= fxi * b;
            jyb -= fdy * (DSy[j] * a);
            const double jz_v = fvz * (a * b + t12 * DSy[j]), rho_v = cd * S1x[i] * S1y[j];
            if (i != ie - 1 && jxb[j] != 0.0) atomicAdd(v.jx + id, jxb[j]);  // (last row / column: see deposit3)
            if (j != je - 1 && jyb != 0.0) atomicAdd(v.jy + id, jyb);
            if (jz_v != 0.0) atomicAdd(v.jz + id, jz_v);
            if (rho_v != 0.0) atomicAdd(v.rho + id, rho_v);
        }
    }
}


inline Slots make_slots(const Species &sp) {
    Slots s;
    s.x = sp.attr[LPIC_P_X]; s.y = sp.attr[LPIC_P_Y]; s.z = sp.attr[LPIC_P_Z]; s.w = sp.attr[LPIC_P_W];
    s.ux = sp.attr[LPIC_P_UX]; s.uy = sp.attr[LPIC_P_UY]; s.uz = sp.attr[LPIC_P_UZ]; s.ig = sp.attr[LPIC_P_INV_GAMMA];
    for (int a = 0; a < 6; a++) s.part[a] = sp.attr[LPIC_P_EX_PART + a];
    s.dead = sp.dead; s.off = sp.d_off; s.npart = sp.d_npart; s.ps = sp.pstride;
    return s;
}
