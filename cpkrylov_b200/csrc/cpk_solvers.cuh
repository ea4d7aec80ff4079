// cpk_solvers.cuh -- the six constraint-preconditioned Krylov loops of cpkrylov,
// each as ONE device function that runs the whole `while` on the GPU: no host
// round trip per iteration, scalars recomputed redundantly (bit-identically) by
// every thread from team-wide deterministic reductions.
//
// Vector convention: every Krylov vector is stored as one N-vector [v; q]
// (n-part, then m-part); H and C are applied together as blkdiag(H, C).
// Statement order and association follow the reference line by line; the file
// is compiled with -fmad=false so a*b+c is rounded twice as in MATLAB.
#pragma once
#include "cpk_kernels.cuh"

namespace cpk {

constexpr double kEps = 2.220446049250313e-16;

template <class Team>
struct Ctx {
    Team &T;
    const DevSystem &S;
    const SolveArgs &A;
    DevStatus *st;
    PhaseClock pc;
    int epoch;
    int hop_seq;            // requests made to a matrix-free A so far (DevHostOp)
    double *dsm;            // dynamic shared memory (doubles)
    int n, m, N;

    __device__ double *vec(int i) const { return A.work + (size_t)i * N; }
    // matrix-free A only: false once the wait for the host's A*v gave up (host error or timeout),
    // so that the loops end instead of iterating on garbage; explicit matrices never look
    __device__ bool alive() const { return !S.hop.on || !T.aborted(); }
    __device__ void hist(int row, long long idx, double v) const {
        if (T.leader() && A.hist && idx < A.hist_cap) A.hist[(size_t)row * A.hist_cap + idx] = v;
    }
    __device__ void fail(int code, int iter, double val, int second = 0) const {
        if (T.leader()) { st->err = code; st->err_iter = iter; st->err_value = val; st->err_second = second; }
    }
    // U = blkdiag(H,C) * V with alpha = U'V fused (e.g. cpminres.m:187-189)
    __device__ double spmv_dot(const double *V, double *U) {
        pc.mark(CPK_PH_VEC_);
        double part[1] = {0.0};
        kkt_mult(T, S, hop_seq, V, [&](int row, double s) { U[row] = s; part[0] += s * V[row]; });
        T.template reduce<1>(part);
        pc.mark(CPK_PH_SPMV_);
        return part[0];
    }
    __device__ void apply(const double *z, bool neg_tail, double *y) {
        VecIn in{z, nullptr, n, neg_tail};
        ldl2_apply(T, S.M, in, y, epoch, st, pc);
    }
    template <class Rider>
    __device__ bool apply_with(const VecIn &in, double *y, Rider rider, double *sum) {
        return ldl2_apply(T, S.M, in, y, epoch, st, pc, rider, sum);
    }
};

__device__ __forceinline__ double dsign(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0); }

// util/SymGivens.m:1-29
__device__ __forceinline__ void sym_givens(double a, double b, double &c, double &s, double &d)
{
    if (b == 0.0) {
        c = (a == 0.0) ? 1.0 : dsign(a);
        s = 0.0;
        d = fabs(a);
    } else if (a == 0.0) {
        c = 0.0;
        s = dsign(b);
        d = fabs(b);
    } else if (fabs(b) > fabs(a)) {
        const double t = a / b;
        s = dsign(b) / sqrt(1.0 + t * t);
        c = s * t;
        d = b / s;
    } else {
        const double t = b / a;
        c = dsign(a) / sqrt(1.0 + t * t);
        s = c * t;
        d = a / c;
    }
}

// g'r + t'w of cpcg.m:167-168 as a rider on the apply's residual pass: row < n
// contributes g*r, row >= n contributes (a + u)*w  ([g;w] is the apply's input, [r;u] its output)
struct CgRider {
    static constexpr bool kActive = true;
    const double *X, *PQ; double alpha; int n;          // a = X + alpha*PQ is still pending (it lands in memory with the p,q update)
    struct Pre { double x, pq; };
    __device__ __forceinline__ Pre pre(int row) const {
        Pre P;
        P.x = row >= n ? X[row] : 0.0;          // loads only, no branch around them
        P.pq = row >= n ? PQ[row] : 0.0;
        return P;
    }
    __device__ __forceinline__ double fin(int row, double gw, double ru, const Pre &P) const {
        return (row < n) ? gw * ru : ((P.x + alpha * P.pq) + ru) * gw;
    }
};

// ---------------------------------------------------------------------------
// kernels/cpcg.m:118-193
// work vectors: 0 X=[x;a] 1 GW=[g;w] 2 PQ=[p;q] 3 APCQ 4 RU
// ---------------------------------------------------------------------------
template <class Team>
__device__ void run_cpcg(Ctx<Team> &c, const double *b, double *X)
{
    Team &T = c.T;
    const int n = c.n, N = c.N;
    double *GW = c.vec(0), *PQ = c.vec(1), *APCQ = c.vec(2), *RU = c.vec(3);
    TEAM_FOR(T, i, N) { X[i] = 0.0; GW[i] = (i < n) ? -b[i] : 0.0; }
    T.sync();
    c.apply(GW, false, RU);                                         // :125
    double part[1] = {0.0};
    TEAM_FOR(T, i, N) { PQ[i] = -RU[i]; if (i < n) part[0] += GW[i] * RU[i]; }
    T.template reduce<1>(part);                                     // also publishes PQ
    double rn2 = part[0];                                           // :130
    if (rn2 < 0.0) { c.fail(CPK_ERR_BREAKDOWN_, 0, rn2); return; }
    double residNorm = sqrt(rn2);
    const double stopTol = c.A.atol + c.A.rtol * residNorm;
    c.hist(0, 0, residNorm);
    long long itn = 0;
    while (residNorm > stopTol && itn < c.A.itmax && c.alive()) {                // :147
        ++itn;
        const double pAp_qCq = c.spmv_dot(PQ, APCQ);                // :151-152
        const double alpha = rn2 / pAp_qCq;                         // :154
        // :161-164.  x,a += alpha*(p,q) is deferred to the p,q update below (same pass).
        // g,w += alpha*(Ap,Cq): with refinement on, the apply evaluates it on the fly and
        // stores it during its residual pass; otherwise it is a pass of its own.
        VecIn in{GW, nullptr, n, false};
        if (c.S.M.nitref > 0) { in.add = APCQ; in.scale = alpha; in.wb = GW; }
        else {
            const double *const src[2] = {GW, APCQ};
            team_map<2>(T, N, src, [&](int i, const double (&v)[2]) { GW[i] = v[0] + alpha * v[1]; });
            T.sync();
        }
        // :166-168.  g'r + t'w rides on the refinement-residual pass of the apply when that
        // pass sees the final [r;u]; otherwise it is a pass of its own.
        double rsum = 0.0;
        const bool fused = c.apply_with(in, RU, CgRider{X, PQ, alpha, n}, &rsum);
        if (fused) part[0] = rsum;
        else {
            part[0] = 0.0;
            const double *const src[4] = {GW, RU, X, PQ};
            team_map<4>(T, N, src, [&](int i, const double (&v)[4]) {
                if (i < n) part[0] += v[0] * v[1];
                else { const double t = (v[2] + alpha * v[3]) + v[1]; part[0] += t * v[0]; }
            });
            T.template reduce<1>(part);
        }
        const double rn2_new = part[0];
        const double beta = rn2_new / rn2;                          // :169
        {                                                           // :161-162 and :171-172
            const double *const src[3] = {RU, PQ, X};       // RU is dead after this pass, X rests until the next one
            team_map<3, 5u>(T, N, src, [&](int i, const double (&v)[3]) {
                const double xi = v[2] + alpha * v[1];
                st_dead(&X[i], xi);
                const double t = (i < n) ? v[0] : xi + v[0];
                PQ[i] = -t + beta * v[1];
            });
        }
        rn2 = rn2_new;
        if (rn2 < 0.0) { c.fail(CPK_ERR_BREAKDOWN_, (int)itn, rn2); break; }
        residNorm = sqrt(rn2);
        c.hist(0, itn, residNorm);
        T.sync();
        c.pc.mark(CPK_PH_VEC_);
    }
    if (T.leader()) {
        c.st->niters = itn; c.st->hist_len = itn + 1;
        c.st->solved = residNorm <= stopTol;
    }
}

// ---------------------------------------------------------------------------
// Lanczos pieces shared by cpcglanczos / cpminres / cpsymmlq
// ---------------------------------------------------------------------------
// v1, q1 and beta1^2 (cpminres.m:131-134): VKP1 = [vprec1; -vprec2], returns u'v1
template <class Team>
__device__ double lanczos_start(Ctx<Team> &c, const double *b, double *U, double *VPREC, double *VKP1)
{
    Team &T = c.T;
    const int n = c.n, N = c.N;
    TEAM_FOR(T, i, N) U[i] = (i < n) ? b[i] : 0.0;
    T.sync();
    c.apply(U, false, VPREC);
    double part[1] = {0.0};
    TEAM_FOR(T, i, N) {
        const double v = (i < n) ? VPREC[i] : -VPREC[i];
        VKP1[i] = v;
        if (i < n) part[0] += U[i] * v;
    }
    T.template reduce<1>(part);
    return part[0];
}

// u = A vk, t = C qk, alpha, vprec = M[u;-t], unnormalised v_{k+1}, q_{k+1} and
// beta^2 = u'v_{k+1} + t'q_{k+1}   (cpminres.m:187-194).  VKM1 == nullptr drops
// the beta*v_{k-1} term (cpsymmlq.m:202-204).
// The Lanczos three-term update and its trailing P-inner product as a rider on the
// apply's residual pass (cpminres.m:191-194): row r of the pass sees the apply's input
// xi = [u; -t]_r and output yi = vprec_r, writes v_{k+1} / q_{k+1} and returns u_r*v_r
// (resp. t_r*q_r).  Association as in the reference: (vprec - alpha*vk) - beta*vkm1 and
// ((qk - vprec) - alpha*qk) - beta*qkm1.
struct LanczosRider {
    static constexpr bool kActive = true;
    const double *VK, *VKM1;        // VKM1 == nullptr: no beta term (cpsymmlq.m:202-204)
    double *VKP1;
    double alpha, beta;
    int n;
    struct Pre { double vk, vkm1; };
    __device__ __forceinline__ Pre pre(int row) const {
        Pre P;
        P.vk = VK[row];
        P.vkm1 = (VKM1 ? VKM1 : VK)[row];       // unused without VKM1
        return P;
    }
    __device__ __forceinline__ double fin(int row, double xi, double yi, const Pre &P) const {
        const double vk = P.vk;
        double v, u;
        if (row < n) { u = xi; v = yi - alpha * vk; }
        else { u = -xi; v = vk - yi; v = v - alpha * vk; }
        if (VKM1) v = v - beta * P.vkm1;
        VKP1[row] = v;
        return u * v;
    }
};

template <class Team>
__device__ void lanczos_step(Ctx<Team> &c, const double *VK, const double *VKM1, double beta,
                             double *U, double *VPREC, double *VKP1, double &alpha, double &betasq)
{
    Team &T = c.T;
    const int n = c.n, N = c.N;
    alpha = c.spmv_dot(VK, U);
    double rsum = 0.0;
    VecIn in{U, nullptr, n, true};
    if (c.apply_with(in, VPREC, LanczosRider{VK, VKM1, VKP1, alpha, beta, n}, &rsum)) { betasq = rsum; return; }
    double part[1] = {0.0};
    if (VKM1) {
        TEAM_FOR(T, i, N) {
            double v;
            if (i < n) v = VPREC[i] - alpha * VK[i] - beta * VKM1[i];
            else { v = VK[i] - VPREC[i]; v = v - alpha * VK[i] - beta * VKM1[i]; }
            VKP1[i] = v;
            part[0] += U[i] * v;
        }
    } else {
        TEAM_FOR(T, i, N) {
            double v;
            if (i < n) v = VPREC[i] - alpha * VK[i];
            else { v = VK[i] - VPREC[i]; v = v - alpha * VK[i]; }
            VKP1[i] = v;
            part[0] += U[i] * v;
        }
    }
    T.template reduce<1>(part);
    betasq = part[0];
}

// ---------------------------------------------------------------------------
// kernels/cpminres.m:114-252
// work: 0 U 1 VPREC 2..4 Lanczos ring 5..7 W ring
// ---------------------------------------------------------------------------
template <class Team>
__device__ void run_cpminres(Ctx<Team> &c, const double *b, double *X)
{
    Team &T = c.T;
    const int n = c.n, N = c.N;
    double *U = c.vec(0), *VPREC = c.vec(1);
    double *VKM1 = c.vec(2), *VK = c.vec(3), *VKP1 = c.vec(4);
    double *WV1 = c.vec(5), *WV2 = c.vec(6), *WV = c.vec(7);
    const double eps100 = 100.0 * kEps;

    double beta = lanczos_start(c, b, U, VPREC, VKP1);
    if (beta < -eps100) { c.fail(CPK_ERR_INDEFINITE_, 0, beta); return; }
    beta = sqrt(fabs(beta));
    TEAM_FOR(T, i, N) {
        X[i] = 0.0; VK[i] = 0.0; WV2[i] = 0.0;
        double v = VKP1[i];
        if (beta > 0.0) v = v / beta;
        VKP1[i] = v; WV[i] = v;
    }
    T.sync();
    double residNorm = beta;
    c.hist(0, 0, residNorm);
    long long k = 0;
    double deltabar = 0.0, epsln = 0.0, taubar = beta, cs = -1.0, sn = 0.0;
    const double stopTol = c.A.atol + c.A.rtol * residNorm;

    while (residNorm > stopTol && k < c.A.itmax && c.alive()) {                  // :176
        ++k;
        { double *t = VKM1; VKM1 = VK; VK = VKP1; VKP1 = t; }       // :181-184
        double alpha, betasq;
        lanczos_step(c, VK, VKM1, beta, U, VPREC, VKP1, alpha, betasq);
        if (betasq < -eps100) { c.fail(CPK_ERR_INDEFINITE_, (int)k, betasq); break; }
        beta = sqrt(fabs(betasq));
        const double oldeps = epsln;                                // :211-215
        const double delta = cs * deltabar + sn * alpha;
        const double gammabar = sn * deltabar - cs * alpha;
        epsln = sn * beta;
        deltabar = -cs * beta;
        const double gamma = hypot(gammabar, beta);                 // :218
        cs = gammabar / gamma;
        sn = beta / gamma;
        const double tau = cs * taubar;
        taubar = sn * taubar;
        { double *t = WV1; WV1 = WV2; WV2 = WV; WV = t; }           // :225-228
        TEAM_FOR(T, i, N) {
            if (beta > 0.0) VKP1[i] = VKP1[i] / beta;               // :203-204
            const double w = (VK[i] - oldeps * WV1[i] - delta * WV2[i]) / gamma;
            WV[i] = w;
            X[i] = (i < n) ? X[i] + tau * w : X[i] - tau * w;       // :231-232
        }
        residNorm = taubar;
        c.hist(0, k, residNorm);
        T.sync();
        c.pc.mark(CPK_PH_VEC_);
    }
    if (T.leader()) { c.st->niters = k; c.st->hist_len = k + 1; c.st->solved = residNorm <= stopTol; }
}

// ---------------------------------------------------------------------------
// kernels/cpcglanczos.m:135-325
// work: 0 U 1 VPREC 2..4 Lanczos ring 5 WV
// ---------------------------------------------------------------------------
template <class Team>
__device__ void run_cpcglanczos(Ctx<Team> &c, const double *b, double *X)
{
    Team &T = c.T;
    const int n = c.n, N = c.N;
    double *U = c.vec(0), *VPREC = c.vec(1);
    double *VKM1 = c.vec(2), *VK = c.vec(3), *VKP1 = c.vec(4), *WV = c.vec(5);
    const double eps100 = 100.0 * kEps;
    const double btol = c.A.btol;

    double beta = lanczos_start(c, b, U, VPREC, VKP1);
    if (beta < -eps100) { c.fail(CPK_ERR_INDEFINITE_, 0, beta); return; }
    beta = sqrt(fabs(beta));
    TEAM_FOR(T, i, N) {
        X[i] = 0.0; VK[i] = 0.0;
        double v = VKP1[i];
        if (beta > 0.0) v = v / beta;
        VKP1[i] = v; WV[i] = v;
    }
    T.sync();
    const double beta1 = beta;
    double residNorm = beta1;
    c.hist(0, 0, residNorm);
    long long k = 0;
    double dg = 0.0, low = 1.0, eta = beta, oldbeta = 0.0, opNorm2 = 0.0;
    double rhobar = 1.0, xxNorm2 = 0.0, xNorm = 0.0, tau = 0.0, delta = 0.0;
    const double stopTol = c.A.atol + c.A.rtol * residNorm;
    double bstopTol = btol * beta1;

    while (residNorm > stopTol && residNorm > bstopTol && k < c.A.itmax && c.alive()) {      // :221
        ++k;
        { double *t = VKM1; VKM1 = VK; VK = VKP1; VKP1 = t; }
        double alpha, betasq;
        lanczos_step(c, VK, VKM1, beta, U, VPREC, VKP1, alpha, betasq);
        dg = alpha - low * low * dg;                                            // :236
        const double zeta = eta / dg;
        if (betasq < -eps100) { c.fail(CPK_ERR_INDEFINITE_, (int)k, betasq); break; }
        beta = sqrt(fabs(betasq));
        low = beta / dg;                                                        // :265
        eta = -low * eta;
        TEAM_FOR(T, i, N) {
            const double w = WV[i];
            X[i] = (i < n) ? X[i] + zeta * w : X[i] - zeta * w;                 // :238-239
            double v = VKP1[i];
            if (beta > 0.0) { v = v / beta; VKP1[i] = v; }                      // :258-259
            WV[i] = v - low * w;                                                // :267-268
        }
        if (btol > 0.0) {                                                       // :271-291
            const double rho = sqrt(rhobar * rhobar + low * low);
            const double cs = rhobar / rho;
            const double sn = low / rho;
            const double num = zeta - delta * tau;
            const double taub = num / rhobar;
            tau = num / rho;
            xNorm = sqrt(xxNorm2 + taub * taub);
            xxNorm2 = xxNorm2 + tau * tau;
            delta = sn;
            rhobar = -cs;
            opNorm2 = opNorm2 + alpha * alpha + beta * beta + oldbeta * oldbeta;
            const double opNorm = sqrt(opNorm2);
            const double bkerr = opNorm * xNorm + beta1;
            bstopTol = btol * bkerr;
        }
        residNorm = beta * fabs(zeta);                                          // :293
        c.hist(0, k, residNorm);
        oldbeta = beta;
        T.sync();
        c.pc.mark(CPK_PH_VEC_);
    }
    if (T.leader()) {
        c.st->niters = k; c.st->hist_len = k + 1;
        int solved = 0, status = 0;
        if (residNorm <= stopTol) { solved = 1; status = 1; }
        if (btol > 0.0 && residNorm <= bstopTol) { solved = 1; status = 2; }
        c.st->solved = solved; c.st->status = status;
    }
}

// ---------------------------------------------------------------------------
// kernels/cpsymmlq.m:121-367
// work: 0 U 1 VPREC 2..4 Lanczos ring 5 WV
// hist rows: 0 cg, 1 lq, 2 qr
// ---------------------------------------------------------------------------
template <class Team>
__device__ void run_cpsymmlq(Ctx<Team> &c, const double *b, double *X)
{
    Team &T = c.T;
    const int n = c.n, N = c.N;
    double *U = c.vec(0), *VPREC = c.vec(1);
    double *VKM1 = c.vec(2), *VK = c.vec(3), *VKP1 = c.vec(4), *WV = c.vec(5);
    const double eps100 = 100.0 * kEps;

    double beta1 = lanczos_start(c, b, U, VPREC, VKP1);
    if (beta1 < -eps100) { c.fail(CPK_ERR_INDEFINITE_, 0, beta1); return; }
    beta1 = sqrt(fabs(beta1));
    TEAM_FOR(T, i, N) {
        X[i] = 0.0; WV[i] = 0.0;
        if (beta1 > 0.0) VKP1[i] = VKP1[i] / beta1;
    }
    T.sync();
    double cgresidNorm = beta1;
    const double stopTol = c.A.atol + c.A.rtol * cgresidNorm;
    long long k = 0;
    long long hlen;
    if (cgresidNorm <= stopTol) {                                               // :161-166
        c.hist(0, 0, beta1); c.hist(1, 0, beta1); c.hist(2, 0, beta1);
        hlen = 1;
    } else {
        { double *t = VK; VK = VKP1; VKP1 = t; }                                // :194-195
        double alpha, betasq;
        lanczos_step(c, VK, (const double *)nullptr, 0.0, U, VPREC, VKP1, alpha, betasq);
        if (betasq < -eps100) { c.fail(CPK_ERR_INDEFINITE_, 0, betasq, 1); return; }
        double beta = sqrt(fabs(betasq));
        if (beta > 0.0) { TEAM_FOR(T, i, N) VKP1[i] = VKP1[i] / beta; }
        T.sync();
        double gammabar = alpha, deltabar = beta, epsdelzeta = beta1, epsilonzeta = 0.0;
        double bstep = 0.0, snprod = 1.0, matnorm2 = alpha * alpha + beta * beta;
        double lqresidNorm, qrresidNorm, den;
        c.hist(0, 0, beta1);                                                    // :331 (prepended)
        long long h = 0;                                                        // entries in lq/qr rows
        while (cgresidNorm > stopTol && k < c.A.itmax && c.alive()) {                        // :229
            double matnorm = sqrt(matnorm2);
            double epsmat = matnorm * kEps;
            den = gammabar;
            if (den == 0.0) den = epsmat;
            lqresidNorm = hypot(epsdelzeta, epsilonzeta);
            qrresidNorm = snprod * beta1;
            cgresidNorm = qrresidNorm * beta / fabs(den);
            c.hist(1, h, lqresidNorm); c.hist(2, h, qrresidNorm); c.hist(0, h + 1, cgresidNorm);
            ++h;
            ++k;
            { double *t = VKM1; VKM1 = VK; VK = VKP1; VKP1 = t; }
            const double betaold = beta;
            lanczos_step(c, VK, VKM1, beta, U, VPREC, VKP1, alpha, betasq);
            if (betasq < -eps100) { c.fail(CPK_ERR_INDEFINITE_, (int)k, betasq); return; }
            beta = sqrt(fabs(betasq));
            matnorm2 = matnorm2 + alpha * alpha + beta * beta + betaold * betaold;
            const double gamma = hypot(gammabar, betaold);                      // :291-297
            const double cs = gammabar / gamma;
            const double sn = betaold / gamma;
            const double delta = cs * deltabar + sn * alpha;
            gammabar = sn * deltabar - cs * alpha;
            const double epsilon = sn * beta;
            deltabar = -cs * beta;
            const double zeta = epsdelzeta / gamma;                             // :300-306
            const double zcs = zeta * cs;
            const double zsn = zeta * sn;
            TEAM_FOR(T, i, N) {
                if (beta > 0.0) VKP1[i] = VKP1[i] / beta;
                const double w = WV[i], v = VK[i];
                X[i] = (i < n) ? X[i] + zcs * w + zsn * v : X[i] - zcs * w - zsn * v;
                WV[i] = sn * w - cs * v;
            }
            bstep = bstep + snprod * cs * zeta;                                 // :310-313
            snprod = snprod * sn;
            epsdelzeta = epsilonzeta - delta * zeta;
            epsilonzeta = -epsilon * zeta;
            T.sync();
            c.pc.mark(CPK_PH_VEC_);
        }
        {                                                                       // :318-327
            double matnorm = sqrt(matnorm2);
            double epsmat = matnorm * kEps;
            den = gammabar;
            if (den == 0.0) den = epsmat;
            lqresidNorm = hypot(epsdelzeta, epsilonzeta);
            qrresidNorm = snprod * beta1;
            c.hist(1, h, lqresidNorm); c.hist(2, h, qrresidNorm);
            ++h;
        }
        hlen = h;
        double zetabar = 0.0;
        const bool to_cg = cgresidNorm < lqresidNorm;                           // :334
        if (to_cg) { zetabar = epsdelzeta / den; bstep = bstep + snprod * zetabar; }
        // :342-347  one more apply, M*[b;0]
        TEAM_FOR(T, i, N) U[i] = (i < n) ? b[i] : 0.0;
        T.sync();
        c.apply(U, false, VPREC);
        bstep = bstep / beta1;
        TEAM_FOR(T, i, N) {
            double xi = X[i];
            if (to_cg) xi = (i < n) ? xi + zetabar * WV[i] : xi - zetabar * WV[i];
            // vk = vprec1, qk = -vprec2:  x + bstep*vk ;  y - bstep*qk
            xi = (i < n) ? xi + bstep * VPREC[i] : xi - bstep * (-VPREC[i]);
            X[i] = xi;
        }
    }
    if (T.leader()) { c.st->niters = k; c.st->hist_len = hlen; c.st->solved = cgresidNorm <= stopTol; }
}

// ---------------------------------------------------------------------------
// v_i = init(i) - sum_j cf[j] * B[cols[j]][i], subtracted in j order (the order of the
// reference's loops cpgmres.m:214-218, cpdqgmres.m:210-216 / :255-258), then fin(i, v_i).
// The pass streams nc columns of the basis: two consecutive elements per thread (16-byte
// loads) and the loads of four columns issued before their subtractions, so that a thread
// has 64 bytes per vector group in flight instead of one 8-byte load per loop trip.
// cols / cf live in shared memory.
// ---------------------------------------------------------------------------
// A pass over the basis streams nc x 8N bytes (cfg 4, mem = 20: 400 MB) past an L2 that should keep the
// matrix and the work vectors of the next phase: columns are loaded evict-first (CPK_BASIS_HINT).
#ifndef CPK_BASIS_HINT
#define CPK_BASIS_HINT 1
#endif
__device__ __forceinline__ double2 ld_basis(const double2 *p) { return CPK_BASIS_HINT ? __ldcs(p) : *p; }
template <class Team, class Init, class Fin>
__device__ __forceinline__ void basis_combine(const Team &T, int N, const double *B, const int *cols, const double *cf, int nc,
                                              Init &&init, Fin &&fin)
{
    const int tid = T.tid, nth = T.nthreads;
    if ((N & 1) == 0 && (reinterpret_cast<size_t>(B) & 15) == 0) {
        const int NP = N >> 1;
        for (int p = tid; p < NP; p += nth) {
            const int i = 2 * p;
            double v0 = init(i), v1 = init(i + 1);
            int j = 0;
            for (; j + 4 <= nc; j += 4) {
                double2 a[4]; double h[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    a[q] = ld_basis(reinterpret_cast<const double2 *>(B + (size_t)cols[j + q] * N + i));
                    h[q] = cf[j + q];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) { v0 = v0 - h[q] * a[q].x; v1 = v1 - h[q] * a[q].y; }
            }
            for (; j < nc; ++j) {
                const double2 a = ld_basis(reinterpret_cast<const double2 *>(B + (size_t)cols[j] * N + i));
                const double h = cf[j];
                v0 = v0 - h * a.x; v1 = v1 - h * a.y;
            }
            fin(i, v0); fin(i + 1, v1);
        }
    } else {
        for (int i = tid; i < N; i += nth) {
            double v = init(i);
            for (int j = 0; j < nc; ++j) v = v - cf[j] * B[(size_t)cols[j] * N + i];
            fin(i, v);
        }
    }
}

// ---------------------------------------------------------------------------
// Arnoldi helpers for cpgmres / cpdqgmres.  The reference's Gram-Schmidt
// coefficients use the FIXED u, t (cpgmres.m:215, cpdqgmres.m:213), so all of
// them come out of one pass over the basis; the subtraction then runs in the
// reference's j order.
//   cols[j] (j < nc) = column index into the basis VQ for the j-th coefficient
//   hs (shared)      = coefficients out
// ---------------------------------------------------------------------------
template <class Team>
__device__ void multi_dot(Ctx<Team> &c, const double *VQ, const int *cols, int nc, const double *U, double *hs)
{
    Team &T = c.T;
    const int N = c.N;
    if (Team::kKind == 1 && N <= 8192) {
        // one-CTA team, short vectors: a warp per column (all loads of a column independent,
        // one butterfly, no CTA barrier per chunk of columns)
        __syncthreads();                    // hs is free (previous coefficients consumed)
        for (int j = T.gwarp; j < nc; j += T.nwarps) {
            const double *col = VQ + (size_t)cols[j] * N;
            double acc = 0.0;
            for (int i0 = T.lane; i0 < N; i0 += 128) {
                double a[4], u[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { const int i = i0 + 32 * q; a[q] = (i < N) ? col[i] : 0.0; u[q] = (i < N) ? U[i] : 0.0; }
#pragma unroll
                for (int q = 0; q < 4; ++q) acc += a[q] * u[q];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
            if (T.lane == 0) hs[j] = acc;
        }
        __syncthreads();
        return;
    }
    for (int j0 = 0; j0 < nc; j0 += kRedMax) {
        const int nv = min(kRedMax, nc - j0);
        double acc[kRedMax];
        const double *colp[kRedMax];
#pragma unroll
        for (int j = 0; j < kRedMax; ++j) { acc[j] = 0.0; colp[j] = VQ + (size_t)cols[j0 + min(j, nv - 1)] * N; }
        if ((N & 1) == 0 && ((reinterpret_cast<size_t>(VQ) | reinterpret_cast<size_t>(U)) & 15) == 0) {
            // two consecutive elements per thread: 16-byte loads, 128 bytes per thread in flight
            const int NP = N >> 1;
            for (int p = T.tid, stride = T.nthreads; p < NP; p += stride) {
                const double2 u = *reinterpret_cast<const double2 *>(U + 2 * p);
                double2 a[kRedMax];
#pragma unroll
                for (int j = 0; j < kRedMax; ++j) a[j] = ld_basis(reinterpret_cast<const double2 *>(colp[j] + 2 * p));
#pragma unroll
                for (int j = 0; j < kRedMax; ++j) { acc[j] += a[j].x * u.x; acc[j] += a[j].y * u.y; }
            }
        } else {
            TEAM_FOR(T, i, N) {
                const double u = U[i];
#pragma unroll
                for (int j = 0; j < kRedMax; ++j) acc[j] += colp[j][i] * u;
            }
        }
        T.template wide_store<kRedMax>(acc, j0, nv, hs);
    }
    T.wide_collect(nc, hs);
}

// ---------------------------------------------------------------------------
// kernels/cpgmres.m:130-269
// work: 0 U 1 W 2.. basis VQ (restart+1 columns)
// dynamic shared (doubles): hs[r+1] cs[r] sn[r] g[r+1] z[r] ; cols (int) after
// global scratch gs: H (r+1) x r column-major (team leader only)
// ---------------------------------------------------------------------------
template <class Team>
__device__ void run_cpgmres(Ctx<Team> &c, const double *b, double *X)
{
    Team &T = c.T;
    const int n = c.n, N = c.N;
    const int R = c.A.restart;
    double *U = c.vec(0), *W = c.vec(1), *VQ = c.vec(2);
    double *hs = c.dsm, *csn = hs + (R + 1), *snn = csn + R, *g = snn + R, *z = g + (R + 1);
    int *cols = (int *)(z + R);
    double *Hg = c.A.gs;                        // (R+1) x R
    __shared__ double s_resid;

    TEAM_FOR(T, i, N) X[i] = 0.0;
    for (int j = threadIdx.x; j <= R; j += blockDim.x) cols[j] = j;
    bool finished = false;
    long long outer = 0;
    const long long outermax = (c.A.itmax + R - 1) / R;             // :148
    double residNorm = 0.0, stopTol = 0.0;
    long long hl = 0;
    int k = 0;
    T.sync();
    while (!finished && outer < outermax && c.alive()) {                         // :155
        ++outer;
        double part[1] = {0.0};
        if (outer == 1) {                                           // :160-165
            TEAM_FOR(T, i, N) U[i] = (i < n) ? b[i] : 0.0;
            T.sync();
        } else {                                                    // :166-168
            c.pc.mark(CPK_PH_VEC_);
            kkt_mult(T, c.S, c.hop_seq, X, [&](int row, double s) { U[row] = (row < n) ? b[row] - s : s; });
            T.sync();
            c.pc.mark(CPK_PH_SPMV_);
        }
        c.apply(U, true, W);
        double *V1 = VQ;
        TEAM_FOR(T, i, N) {
            double v;
            if (i < n) v = W[i];
            else v = (outer == 1) ? -W[i] : X[i] - W[i];            // :165 / :171
            V1[i] = v;
            part[0] += U[i] * v;
        }
        T.template reduce<1>(part);
        if (part[0] < 0.0) { c.fail(CPK_ERR_BREAKDOWN_, (int)((outer - 1) * R), part[0]); return; }
        residNorm = sqrt(part[0]);                                  // :173
        if (residNorm != 0.0) { TEAM_FOR(T, i, N) V1[i] = V1[i] / residNorm; }
        if (outer == 1) { stopTol = c.A.atol + c.A.rtol * residNorm; c.hist(0, 0, residNorm); hl = 1; }
        k = 0;
        if (T.cta_leader()) g[0] = residNorm;
        T.sync();
        while (residNorm > stopTol && k < R && c.alive()) {                      // :203
            ++k;
            const double *Vk = VQ + (size_t)(k - 1) * N;
            double *Vk1 = VQ + (size_t)k * N;
            c.pc.mark(CPK_PH_VEC_);
            kkt_mult(T, c.S, c.hop_seq, Vk, [&](int row, double s) { U[row] = s; });   // :209-210
            T.sync();
            c.pc.mark(CPK_PH_SPMV_);
            c.apply(U, true, W);                                    // :211
            multi_dot(c, VQ, cols, k, U, hs);                       // :215
            part[0] = 0.0;
            basis_combine(T, N, VQ, cols, hs, k,                    // :212-218
                          [&](int i) { return (i < n) ? W[i] : Vk[i] - W[i]; },
                          [&](int i, double v) { Vk1[i] = v; part[0] += U[i] * v; });
            T.template reduce<1>(part);
            if (part[0] < 0.0) { c.fail(CPK_ERR_BREAKDOWN_, (int)((outer - 1) * R + k), part[0]); return; }
            const double hk1 = sqrt(part[0]);                       // :219
            if (hk1 != 0.0) { TEAM_FOR(T, i, N) Vk1[i] = Vk1[i] / hk1; }
            // scalar part, one thread per CTA, identical in every CTA (:229-248)
            if (T.cta_leader()) {
                for (int j = 0; j < k - 1; ++j) {
                    const double Hjk = csn[j] * hs[j] + snn[j] * hs[j + 1];
                    hs[j + 1] = snn[j] * hs[j] - csn[j] * hs[j + 1];
                    hs[j] = Hjk;
                }
                double cc, ss, dd;
                sym_givens(hs[k - 1], hk1, cc, ss, dd);
                csn[k - 1] = cc; snn[k - 1] = ss; hs[k - 1] = dd;
                g[k] = ss * g[k - 1];
                g[k - 1] = cc * g[k - 1];
                s_resid = fabs(g[k]);
                if (T.leader()) for (int j = 0; j < k; ++j) Hg[(size_t)(k - 1) * (R + 1) + j] = hs[j];
            }
            T.sync();
            residNorm = s_resid;
            c.hist(0, hl, residNorm); ++hl;
            c.pc.mark(CPK_PH_VEC_);
        }
        // z = H(1:k,1:k) \ g(1:k) by the team leader's warp (:257), then x += V z, y -= Q z
        if (T.gwarp == 0) {
            for (int i = k - 1; i >= 0; --i) {
                double acc = 0.0;
                for (int j = i + 1 + T.lane; j < k; j += 32) acc += Hg[(size_t)j * (R + 1) + i] * z[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
                if (T.lane == 0) {
                    const double zi = (g[i] - acc) / Hg[(size_t)i * (R + 1) + i];
                    z[i] = zi;
                    c.A.gs[(size_t)(R + 1) * R + i] = zi;
                }
                __syncwarp();
            }
        }
        T.sync();
        for (int j = threadIdx.x; j < k; j += blockDim.x) z[j] = ld_cg(&c.A.gs[(size_t)(R + 1) * R + j]);
        T.cta_sync();
        TEAM_FOR(T, i, N) {                                         // :258-260
            double acc = 0.0;
            for (int j = 0; j < k; ++j) acc += VQ[(size_t)j * N + i] * z[j];
            X[i] = (i < n) ? X[i] + acc : X[i] - acc;
        }
        finished = residNorm <= stopTol;
        T.sync();
        c.pc.mark(CPK_PH_VEC_);
    }
    if (T.leader()) {
        c.st->niters = (outer > 0 ? (outer - 1) * R : 0) + k;       // :267
        c.st->hist_len = hl;
        c.st->solved = residNorm <= stopTol;
    }
}

// ---------------------------------------------------------------------------
// kernels/cpdqgmres.m:125-280
// work: 0 U 1 W 2.. VQ ring (mem+1 columns), then PVQ ring (mem+1 columns)
// dynamic shared (doubles): hs[mem+1] cs[mem] sn[mem] g[mem+1] pc[mem+1]
//                           Hb[(mem+2) x (mem+3)]  band ring of H ; cols (int)
// The reference's H(j, 2+k-j) (itmax x (mem+2), cpdqgmres.m:133) is kept as a
// ring over the live rows j mod (mem+2).
// ---------------------------------------------------------------------------
template <class Team>
__device__ void run_cpdqgmres(Ctx<Team> &c, const double *b, double *X)
{
    Team &T = c.T;
    const int n = c.n, N = c.N;
    const int mem = c.A.mem;
    const int M1 = mem + 1, HR = mem + 2, HC_ = mem + 3;
    double *U = c.vec(0), *W = c.vec(1), *VQ = c.vec(2), *PVQ = c.vec(2 + M1);
    double *hs = c.dsm, *csn = hs + M1, *snn = csn + mem, *g = snn + mem, *pcf = g + M1, *Hb = pcf + M1;
    int *cols = (int *)(Hb + (size_t)HR * HC_);
    __shared__ double s_scal[4];        // residNorm, H(k,2), g(kpos)

#define HB(j, kk) Hb[(size_t)((j) % HR) * HC_ + (kk)]
    for (int j = threadIdx.x; j < HR * HC_; j += blockDim.x) Hb[j] = 0.0;
    TEAM_FOR(T, i, N) { X[i] = 0.0; U[i] = (i < n) ? b[i] : 0.0; }
    T.sync();
    c.apply(U, false, W);                                           // :151
    double part[1] = {0.0};
    TEAM_FOR(T, i, N) {
        const double v = (i < n) ? W[i] : -W[i];
        VQ[i] = v;
        if (i < n) part[0] += U[i] * v;
    }
    T.template reduce<1>(part);
    if (part[0] < 0.0) { c.fail(CPK_ERR_BREAKDOWN_, 0, part[0]); return; }
    double residNorm = sqrt(part[0]);                               // :154
    if (residNorm != 0.0) { TEAM_FOR(T, i, N) VQ[i] = VQ[i] / residNorm; }
    long long k = 0;
    if (T.cta_leader()) g[0] = residNorm;
    const double stopTol = c.A.atol + c.A.rtol * residNorm;
    c.hist(0, 0, residNorm);
    T.sync();
    while (residNorm > stopTol && k < c.A.itmax && c.alive()) {                  // :194
        ++k;
        const int kpos = (int)((k - 1) % M1);                       // 0-based slots (:199-201)
        const int kp1pos = (int)(k % M1);
        const int rotpos = (int)((k - 1) % mem);
        const double *Vk = VQ + (size_t)kpos * N;
        double *Vk1 = VQ + (size_t)kp1pos * N;
        c.pc.mark(CPK_PH_VEC_);
        kkt_mult(T, c.S, c.hop_seq, Vk, [&](int row, double s) { U[row] = s; });       // :205-206
        T.sync();
        c.pc.mark(CPK_PH_SPMV_);
        c.apply(U, true, W);                                        // :207
        const long long jlo = (k - mem + 1 > 1) ? k - mem + 1 : 1;  // :210
        const int nc = (int)(k - jlo + 1);
        T.cta_sync();
        for (int j = threadIdx.x; j < nc; j += blockDim.x) cols[j] = (int)((jlo + j - 1) % M1);
        T.cta_sync();
        multi_dot(c, VQ, cols, nc, U, hs);                          // :213
        part[0] = 0.0;
        basis_combine(T, N, VQ, cols, hs, nc,                       // :208-216
                      [&](int i) { return (i < n) ? W[i] : Vk[i] - W[i]; },
                      [&](int i, double v) { Vk1[i] = v; part[0] += U[i] * v; });
        T.template reduce<1>(part);
        if (part[0] < 0.0) { c.fail(CPK_ERR_BREAKDOWN_, (int)k, part[0]); return; }
        const double hk1 = sqrt(part[0]);                           // :218
        const long long plo = (k - mem > 1) ? k - mem : 1;          // :228, :255
        const int np = (int)(k - plo);
        if (T.cta_leader()) {
            for (int j = 0; j < nc; ++j) HB(jlo + j, 2 + k - (jlo + j)) = hs[j];
            HB(k, 1) = hk1;
            for (long long j = plo; j <= k - 1; ++j) {              // :228-235
                const int jr = (int)((j - 1) % mem);
                const int kk = (int)(k - j + 1), kk1 = kk + 1;
                const double a = HB(j, kk1), bb = HB(j + 1, kk);
                const double Hjk = csn[jr] * a + snn[jr] * bb;
                HB(j + 1, kk) = snn[jr] * a - csn[jr] * bb;
                HB(j, kk1) = Hjk;
            }
            double cc, ss, dd;
            sym_givens(HB(k, 2), HB(k, 1), cc, ss, dd);             // :243
            csn[rotpos] = cc; snn[rotpos] = ss;
            HB(k, 2) = dd; HB(k, 1) = 0.0;
            g[kp1pos] = ss * g[kpos];
            g[kpos] = cc * g[kpos];
            for (int j = 0; j < np; ++j) pcf[j] = HB(plo + j, 2 + k - (plo + j));
            s_scal[0] = fabs(g[kp1pos]);
            s_scal[1] = dd;
            s_scal[2] = g[kpos];
            // the row that leaves the band must read as zero when its slot is reused
            for (int kk = 0; kk < HC_; ++kk) HB(k + 2, kk) = 0.0;
        }
        T.cta_sync();
        for (int j = threadIdx.x; j < np; j += blockDim.x) cols[j] = (int)((plo + j - 1) % M1);
        T.cta_sync();
        const double hk2 = s_scal[1], gk = s_scal[2];
        double *PVk = PVQ + (size_t)kpos * N;
        basis_combine(T, N, PVQ, cols, pcf, np,                     // :253-263
                      [&](int i) {
                          if (hk1 != 0.0) Vk1[i] = Vk1[i] / hk1;    // :222-225
                          return Vk[i];
                      },
                      [&](int i, double p) {
                          p = p / hk2;
                          PVk[i] = p;
                          X[i] = (i < n) ? X[i] + gk * p : X[i] - gk * p;   // :264-265
                      });
        residNorm = s_scal[0];                                      // :268
        c.hist(0, k, residNorm);
        T.sync();
        c.pc.mark(CPK_PH_VEC_);
    }
#undef HB
    if (T.leader()) { c.st->niters = k; c.st->hist_len = k + 1; c.st->solved = residNorm <= stopTol; }
}

// ---------------------------------------------------------------------------
// Entry shared by all launches: optional reg_cpkrylov shift, solver dispatch,
// un-shift (reg_cpkrylov.m:153-173).  X = A.x receives [x; y].
// work vectors N..: the solvers use c.vec(i) above the two reserved here.
// ---------------------------------------------------------------------------
template <int SOLVER, class Team>
__device__ void solve_entry(Team &T, const DevSystem &S, const SolveArgs &A0, double *dsm)
{
    SolveArgs A = A0;
    Ctx<Team> c{T, S, A, A.status, PhaseClock(), 0, 0, dsm, S.n, S.m, S.N};
    c.epoch = *S.M.epoch;       // flags of earlier launches carry earlier epochs
    compact_init(T, S.M);
    const int n = S.n, N = S.N;
    c.pc.start(A.profile && T.leader(), A.status->phase_cycles);
    // two reserved N-vectors at the end of the workspace: XY0 and B1
    double *XY0 = A.work + A.work_len - (size_t)2 * N;
    double *B1 = XY0 + N;
    const double *b1 = A.b;
    bool shift = false;
    if (A.reg_mode) {
        double cnt[1] = {0.0};
        TEAM_FOR(T, i, N) if (i >= n && A.b[i] != 0.0) cnt[0] += 1.0;       // any(b(n+1:n+m)), :154
        T.template reduce<1>(cnt);
        shift = cnt[0] > 0.0;
        if (shift) {
            TEAM_FOR(T, i, N) B1[i] = (i < n) ? 0.0 : A.b[i];
            T.sync();
            VecIn in{B1, nullptr, n, false};
            ldl2_apply(T, S.M, in, XY0, c.epoch, c.st, c.pc);              // :156
            // b1 = b(1:n) - A*xy0(1:n) - B'*xy0(n+1:n+m)                    :157
            auto shift_epi = [&](int row, double s) { B1[row] = A.b[row] - s; };
            if (S.hop.on) hostop_mult(T, S, c.hop_seq, XY0, false, shift_epi);
            else spmv_sell(T, S.Hn, XY0, shift_epi);
            T.sync();
            spmv_sell(T, S.M.K12, XY0 + n, [&](int row, double s) { B1[row] = B1[row] - s; });
            T.sync();
            b1 = B1;
        }
        if (T.leader()) A.status->shifted = shift;
        c.pc.mark(CPK_PH_OTHER_);
    }
    double *X = A.x;
    if (SOLVER == 0) run_cpcg(c, b1, X);
    else if (SOLVER == 1) run_cpcglanczos(c, b1, X);
    else if (SOLVER == 2) run_cpminres(c, b1, X);
    else if (SOLVER == 3) run_cpsymmlq(c, b1, X);
    else if (SOLVER == 4) run_cpgmres(c, b1, X);
    else if (SOLVER == 5) run_cpdqgmres(c, b1, X);
    c.pc.mark(CPK_PH_VEC_);
    if (shift) {                                                            // :166-168
        TEAM_FOR(T, i, N) X[i] = XY0[i] + X[i];
    }
    T.sync();
    compact_drain(T, S.M);
    c.pc.mark(CPK_PH_OTHER_);
    if (T.leader()) {
        *S.M.epoch = c.epoch;
        if (T.aborted()) A.status->err = CPK_ERR_TIMEOUT_;
    }
}

// One persistent kernel per (solver, team kind).  GRID: cooperative launch, the
// whole grid works on sys[0]; otherwise CTA b works on sys[b] (batch).
template <int SOLVER, bool GRID>
__global__ void __launch_bounds__(kBlock, kCtasPerSm)
k_solve(const DevSystem *sys, const SolveArgs *args, TeamCtl *ctl, double *partials, double *wide, int wide_cols, int team_ctas)
{
    __shared__ TeamShared sh;
    // the operator descriptors are read all the time: keep a CTA-local copy in
    // shared memory (global copies would miss in L1 after every barrier)
    __shared__ DevSystem s_sys;
    __shared__ SolveArgs s_args;
    {
        // GRID: the grid is one team, or (team_ctas > 0) consecutive sub-teams of team_ctas CTAs, team t on sys[t]
        const int idx = GRID ? (team_ctas > 0 ? (int)blockIdx.x / team_ctas : 0) : (int)blockIdx.x;
        const int *src = reinterpret_cast<const int *>(sys + idx);
        int *dst = reinterpret_cast<int *>(&s_sys);
        for (int i = threadIdx.x; i < (int)(sizeof(DevSystem) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
        src = reinterpret_cast<const int *>(args + idx);
        dst = reinterpret_cast<int *>(&s_args);
        for (int i = threadIdx.x; i < (int)(sizeof(SolveArgs) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
        if (threadIdx.x == 0) s_sys.M.cw.smem_off = GRID ? -1 : s_args.cw_off;     // per launch: depends on the solver's scratch
        __syncthreads();
    }
    if (GRID) {
        GridTeam T;
        T.init(ctl + (team_ctas > 0 ? (int)blockIdx.x / team_ctas : 0), partials, &sh, team_ctas);
        T.wide = wide; T.wide_cols = wide_cols;
        solve_entry<SOLVER>(T, s_sys, s_args, g_dsm);
    } else {
        CtaTeam T;
        T.init(ctl + blockIdx.x, nullptr, &sh);
        solve_entry<SOLVER>(T, s_sys, s_args, g_dsm);
    }
}

}  // namespace cpk

// Each solver is compiled in its own translation unit (cpk_solve_<name>.cu):
//   CPK_DEFINE_SOLVER_TU(id, name) exports  const void *cpk_kernel_<name>(int grid)
#define CPK_DEFINE_SOLVER_TU(ID, NAME)                                                     \
    extern "C" const void *cpk_kernel_##NAME(int grid)                                     \
    {                                                                                      \
        return grid ? (const void *)cpk::k_solve<ID, true> : (const void *)cpk::k_solve<ID, false>; \
    }
