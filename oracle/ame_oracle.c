/*
 * ame_oracle.c -- CPU restatement of the reference's affine-ME kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under vvc-affine-gpu_b200/ may include,
 * link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker /
 * reported baseline.
 *
 * Parity pin: the restatement is checked against logs produced by the
 * UNMODIFIED reference (oracle/_ref, built by oracle/Makefile from
 * /root/reference/main.cpp) run through NVIDIA's OpenCL on a B200; those logs
 * are committed under tests/golden/ (see tests/golden/README.md).
 *
 * The structure deliberately mirrors the reference's work-group program
 * (one "work-group" = 256 items = all CUs of one size group inside one CTU;
 * barrier-separated phases become loops over lid), NOT the product's per-CU
 * formulation, so that the two are independent derivations of the same result.
 *
 *   aligned kernel      /root/reference/affine.cl:11-958
 *   half-aligned kernel /root/reference/affine.cl:960-1950
 *   helpers             /root/reference/aux_functions.cl (cited per function)
 *
 * Compile with -ffp-contract=off: every fused operation is written explicitly.
 */
#include "ame_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ame_oracle_tables.h"
#ifdef _OPENMP
#include <omp.h>
#endif

#define CTU 128
#define WG 256

typedef struct { int x, y; } mv_t;

/* ---------------------------------------------------------------- helpers */

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int shl(int v, int s) { return (int)((unsigned)v << s); } /* OpenCL '<<' on negatives */
static inline int ilog2(int v) { int r = 0; while (v > 1) { v >>= 1; r++; } return r; }

/* aux_functions.cl:38-47  roundMv */
static mv_t round_mv(mv_t m, int shift) {
    int off = 1 << (shift - 1);
    mv_t r = {(m.x + off - (m.x >= 0)) >> shift, (m.y + off - (m.y >= 0)) >> shift};
    return r;
}

/* aux_functions.cl:51-67  clipMv (block_x/y is the CU origin) */
static mv_t clip_mv(mv_t m, int bx, int by, int W, int H) {
    int horMax = shl(W + 8 - bx - 1, 4), horMin = shl(-128 - 8 - bx + 1, 4);
    int verMax = shl(H + 8 - by - 1, 4), verMin = shl(-128 - 8 - by + 1, 4);
    mv_t r = {clampi(m.x, horMin, horMax), clampi(m.y, verMin, verMax)};
    return r;
}

/* aux_functions.cl:106-141  isSubblockVectorSpreadOverLimit, bipred == 0 */
static int spread_over_limit(int a, int b, int c, int d) {
    const int s4 = 4 << 11, tap = 6;
    int w = (4 * a + s4 > 0 ? 4 * a + s4 : 0) - (4 * a + s4 < 0 ? 4 * a + s4 : 0);
    int h = (4 * b > 0 ? 4 * b : 0) - (4 * b < 0 ? 4 * b : 0);
    w = (w >> 11) + tap + 3;
    h = (h >> 11) + tap + 3;
    if (w * h > (tap + 9) * (tap + 5)) return 1;
    w = (4 * c > 0 ? 4 * c : 0) - (4 * c < 0 ? 4 * c : 0);
    h = (4 * d + s4 > 0 ? 4 * d + s4 : 0) - (4 * d + s4 < 0 ? 4 * d + s4 : 0);
    w = (w >> 11) + tap + 3;
    h = (h >> 11) + tap + 3;
    if (w * h > (tap + 5) * (tap + 9)) return 1;
    return 0;
}

/* aux_functions.cl:146-176 (nCP==2) and :181-212 (nCP==3) */
static mv_t derive_sub_mv(const oracle_cpmvs *c, int nCP, int w, int h, int sx, int sy, int *isSpread) {
    const int shift = 7;
    int cx = sx + 2, cy = sy + 2;
    int dHx = shl(c->RTx - c->LTx, shift - ilog2(w));
    int dHy = shl(c->RTy - c->LTy, shift - ilog2(w));
    int dVx, dVy;
    if (nCP == 3) {
        dVx = shl(c->LBx - c->LTx, shift - ilog2(h));
        dVy = shl(c->LBy - c->LTy, shift - ilog2(h));
    } else {
        dVx = -dHy;
        dVy = dHx;
    }
    int sH = shl(c->LTx, shift), sV = shl(c->LTy, shift);
    int spread = spread_over_limit(dHx, dHy, dVx, dVy);
    mv_t m;
    if (spread) {
        m.x = sH + dHx * (w >> 1) + dVx * (h >> 1);
        m.y = sV + dHy * (w >> 1) + dVy * (h >> 1);
    } else {
        m.x = sH + dHx * cx + dVx * cy;
        m.y = sV + dHy * cx + dVy * cy;
    }
    *isSpread = spread;
    return m;
}

/* affine.cl:246-326: 11x11 window gather with the reference's slack/select logic,
 * kept literal (it is equivalent to clamp-to-edge addressing; tests check that). */
static void gather_window(const int16_t *ref, int W, int H, int px, int py, int ix, int iy, int win[121]) {
    const int N = 8;
    int refPos = py * W + px + iy * W + ix;
    refPos -= ((N >> 1) - 1) * W;
    refPos -= (N / 2 - 1);
    int leftSlack = px + ix - (N / 2 - 1);
    int rightSpam = px + ix + (N / 2);
    int rightSlack = W - 1 - rightSpam;
    int topSlack = py + iy - (N / 2 - 1);
    int bottomSpam = py + iy + (N / 2);
    int bottomSlack = H - 1 - bottomSpam;
    for (int row = 0; row < 11; row++) {
        for (int col = 0; col < 11; col++) {
            int lC = !(leftSlack + col >= 0);
            int rC = !(rightSlack - col + 7 >= 0);
            int tC = !(topSlack + row >= 0);
            int bC = !(bottomSlack - row + 7 >= 0);
            int tl = lC && tC, tr = rC && tC, bl = lC && bC, br = rC && bC;
            lC = lC && !(tl + bl);
            rC = rC && !(tr + br);
            tC = tC && !(tl + tr);
            bC = bC && !(bl + br);
            int idx = refPos + row * W + col;
            if (lC) idx = refPos + row * W - leftSlack;
            if (rC) idx = refPos + row * W + 7 + rightSlack;
            if (tC) idx = refPos + (-topSlack) * W + col;
            if (bC) idx = refPos + (7 + bottomSlack) * W + col;
            if (tl) idx = 0;
            if (tr) idx = W - 1;
            if (bl) idx = (H - 1) * W;
            if (br) idx = W * H - 1;
            win[row * 11 + col] = ref[idx];
        }
    }
}

/* aux_functions.cl:1096-1239 with enablePROF == 0 (affine.cl:168): separable
 * 8-tap filter, horizontal first (shift 2, offset -8192<<2), then vertical
 * (shift 10, offset 512 + (8192<<6)) and clip to [0,1023]. */
static void hv_filter(const int win[121], int xFrac, int yFrac, int pred[16]) {
    int tmp[44];
    const int *cf = O_FILTER[xFrac];
    for (int row = 0; row < 11; row++)
        for (int col = 0; col < 4; col++) {
            int sum = 0;
            for (int k = 0; k < 8; k++) sum += win[row * 11 + col + k] * cf[k];
            tmp[row * 4 + col] = (sum + (-8192 * 4)) >> 2;
        }
    cf = O_FILTER[yFrac];
    for (int row = 0; row < 4; row++)
        for (int col = 0; col < 4; col++) {
            int sum = 0;
            for (int k = 0; k < 8; k++) sum += tmp[(row + k) * 4 + col] * cf[k];
            int val = (sum + (1 << 9) + (8192 << 6)) >> 10;
            pred[row * 4 + col] = clampi(val, 0, 1023);
        }
}

/* aux_functions.cl:1940-2043  satd_4x4 (VTM xCalcHADs4x4), restated as the matrix
 * product C = H * D * H^T with the 4x4 Hadamard matrix: the reference's butterfly
 * computes the same 16 coefficients in another order, and only sum|C| and the DC
 * coefficient C[0][0] = sum(D) enter the result. */
static int satd4x4(const int *org, const int *prd) {
    static const int Hm[4][4] = {{1, 1, 1, 1}, {1, 1, -1, -1}, {1, -1, -1, 1}, {1, -1, 1, -1}};
    int D[4][4], T[4][4], satd = 0, dc = 0;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) D[r][c] = org[r * 4 + c] - prd[r * 4 + c];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            T[r][c] = 0;
            for (int k = 0; k < 4; k++) T[r][c] += Hm[r][k] * D[k][c];
        }
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            int v = 0;
            for (int k = 0; k < 4; k++) v += T[r][k] * Hm[c][k];
            if (r == 0 && c == 0) dc = abs(v);
            satd += abs(v);
        }
    satd -= dc;          /* JVET_R0164 mean-scaled SATD: DC term counted at 1/4 */
    satd += dc >> 2;
    return (satd + 1) >> 1;
}

/* aux_functions.cl:2057-2075, internal(1/16) -> quarter precision */
static mv_t to_quarter(mv_t m) {
    mv_t r = {m.x >= 0 ? (m.x + 2 - 1) >> 2 : (m.x + 2) >> 2, m.y >= 0 ? (m.y + 2 - 1) >> 2 : (m.y + 2) >> 2};
    return r;
}

/* aux_functions.cl:2117-2129 */
static int exp_golomb_bits(int value) {
    unsigned len = 1;
    unsigned t = value <= 0 ? (((unsigned)(-value)) << 1) + 1 : (unsigned)(value << 1);
    while (t > 128) { len += 14; t >>= 7; }
    return (int)len + (((int)floor(log2((double)(float)t))) << 1);
}

/* aux_functions.cl:2132-2137, cost_scale = 0, imvShift = 0 */
static int bits_with_pred(mv_t pred, mv_t sel) {
    return exp_golomb_bits(sel.x - pred.x) + exp_golomb_bits(sel.y - pred.y);
}

/* aux_functions.cl:2140-2189 */
static int calc_affine_bits(int nCP, const oracle_cpmvs *c, const oracle_cpmvs *p) {
    int bits = 0;
    mv_t t, pr, se;
    t.x = p->LTx; t.y = p->LTy; pr = to_quarter(t);
    t.x = c->LTx; t.y = c->LTy; se = to_quarter(t);
    bits += bits_with_pred(pr, se);
    t.x = p->RTx + c->LTx - p->LTx; t.y = p->RTy + c->LTy - p->LTy; pr = to_quarter(t);
    t.x = c->RTx; t.y = c->RTy; se = to_quarter(t);
    bits += bits_with_pred(pr, se);
    t.x = p->LBx + c->LTx - p->LTx; t.y = p->LBy + c->LTy - p->LTy; pr = to_quarter(t);
    t.x = c->LBx; t.y = c->LBy; se = to_quarter(t);
    int extra = bits_with_pred(pr, se);
    return nCP == 3 ? bits + extra : bits;
}

/* aux_functions.cl:2219-2221: float multiply, float floor, int result */
static int get_cost(int bitrate, float lambda) {
    float p = lambda * (float)bitrate;
    return (int)floorf(p);
}

/* double -> int conversion of aux_functions.cl:2203-2210 "(int)(...)": C leaves
 * out-of-range / NaN undefined, so the rule is explicit.
 *   ORACLE_CVT_X86   : cvttsd2si  -> INT_MIN for NaN and anything out of range
 *   ORACLE_CVT_NVIDIA: cvt.rzi.s32.f64 -> 0 for NaN, saturating otherwise     */
static int cvt_d2i(double v, int rule) {
    if (v != v) return rule == ORACLE_CVT_NVIDIA ? 0 : (int)0x80000000;
    if (v >= 2147483648.0) return rule == ORACLE_CVT_NVIDIA ? 0x7fffffff : (int)0x80000000;
    if (v <= -2147483649.0) return (int)0x80000000;
    return (int)v; /* truncation toward zero */
}

/* aux_functions.cl:2194-2215: (int)(d*4 + SIGN(d)*0.5) << 2 */
static int scale_delta(double d, int rule) {
    double s = (d >= 0 ? 1 : -1) * 0.5;
    double v = d * 4 + s; /* d*4 is exact, so fused or not gives the same value */
    return shl(cvt_d2i(v, rule), 2);
}

/* affine.cl:783-855: Gaussian elimination with partial pivoting, VTM solveEqual
 * without its early returns. m is [7][7]; rows 1..n, columns 0..n. */
void oracle_solve(double m[7][7], int n, int fused_backsub, double out[6]) {
    for (int k = 0; k < n; k++) out[k] = 0.;
    for (int i = 1; i < n; i++) {
        double temp = fabs(m[i][i - 1]);
        int tempIdx = i;
        for (int j = i + 1; j < n + 1; j++)
            if (fabs(m[j][i - 1]) > temp) { temp = fabs(m[j][i - 1]); tempIdx = j; }
        if (tempIdx != i)
            for (int j = 0; j < n + 1; j++) {
                m[0][j] = m[i][j];
                m[i][j] = m[tempIdx][j];
                m[tempIdx][j] = m[0][j];
            }
        for (int j = i + 1; j < n + 1; j++)
            for (int k = i; k < n + 1; k++) {
                double prod = m[i][k] * m[j][i - 1];
                double quot = prod / m[i][i - 1];
                m[j][k] = m[j][k] - quot;
            }
    }
    out[n - 1] = m[n][n] / m[n][n - 1];
    for (int i = n - 2; i >= 0; i--) {
        if (m[i + 1][i] == 0.) {
            for (int k = 0; k < n; k++) out[k] = 0.;
            break;
        }
        double temp = 0;
        for (int j = i + 1; j < n; j++) {
            if (fused_backsub) temp = fma(m[i + 1][j], out[j], temp);
            else { double p = m[i + 1][j] * out[j]; temp = temp + p; }
        }
        out[i] = (m[i + 1][n] - temp) / m[i + 1][i];
    }
}

/* ------------------------------------------------------ one work-group run */

typedef struct {
    int16_t tile[CTU * CTU];  /* __local predCU_then_error */
    int16_t gx[CTU * CTU];    /* this WG's slice of horizontalGrad */
    int16_t gy[CTU * CTU];    /* this WG's slice of verticalGrad */
    int64_t eq[WG][7][7];     /* this WG's slice of global_pEqualCoeff */
    int64_t satd[WG];         /* local_cumulativeSATD */
    int cur[WG][4][16];       /* currentCU_subBlock per item */
} wg_scratch;

/* Runs one 256-item work-group: group `g` of CTU `ctuIdx`. ha selects the
 * half-aligned kernel. */
static void run_wg(const oracle_opts *o, const int16_t *ref, const int16_t *cur, int W, int H, float lambda,
                   int ctuIdx, int g, int ha, int nCP, int64_t *gBestCost, oracle_cpmvs *gBestCpmvs,
                   const oracle_cpmvs *gPrev, wg_scratch *s) {
    const int cuW = ha ? O_HA_W[g] : O_W[g];
    const int cuH = ha ? O_HA_H[g] : O_H[g];
    const int cusPerCtu = ha ? O_HA_N[g] : (CTU * CTU) / (cuW * cuH);
    const int itemsPerCu = WG / cusPerCtu;
    const int sbCols = cuW / 4;
    const int ctusPerRow = (int)ceilf((float)W / CTU);
    const int ctuX = (ctuIdx % ctusPerRow) * CTU, ctuY = (ctuIdx / ctusPerRow) * CTU;
    const int cuColumnsPerCtu = CTU / cuW;
    const int perCtu = ha ? O_HA_CUS_PER_CTU : O_ALIGNED_CUS_PER_CTU;
    const int base = ctuIdx * perCtu + (ha ? O_HA_STRIDE[g] : O_STRIDE[g]);
    const int n = 2 * nCP; /* affineParaNum */

    int cuXs[64], cuYs[64];
    for (int c = 0; c < cusPerCtu; c++) {
        cuXs[c] = ha ? O_HA_X[g][c] : (c % cuColumnsPerCtu) * cuW;
        cuYs[c] = ha ? O_HA_Y[g][c] : (c / cuColumnsPerCtu) * cuH;
    }
    /* number of sub-blocks each item predicts: 4 in the aligned kernel
     * (affine.cl:207-209), ceil(area*4/16384) in the HA kernel (:1171-1183) */
    const int nPasses = ha ? (int)ceilf(((float)cuW * cuH * cusPerCtu * 4) / (CTU * CTU)) : 4;
    const int stridePerPass = ha ? itemsPerCu : (cuH * 2) / (CTU / cuW);

    oracle_cpmvs predC[64], currC[64], bestC[64];
    int64_t bestCost[64];
    int within[64];

    memset(s->tile, 0, sizeof s->tile); /* the reference leaves this uninitialised */
    memset(s->gx, 0, sizeof s->gx);
    memset(s->gy, 0, sizeof s->gy);
    memset(s->cur, 0, sizeof s->cur);

    for (int c = 0; c < cusPerCtu; c++) {
        oracle_cpmvs p;
        memset(&p, 0, sizeof p);
        if (nCP == 3) { /* affine.cl:62-106 */
            p = gPrev[base + c];
            int sh = 7 + ilog2(cuH) - ilog2(cuW);
            int vx2 = shl(p.LTx, 7) - shl(p.RTy - p.LTy, sh);
            int vy2 = shl(p.LTy, 7) + shl(p.RTx - p.LTx, sh);
            vx2 = (vx2 + 64 - (vx2 >= 0)) >> 7;
            vy2 = (vy2 + 64 - (vy2 >= 0)) >> 7;
            mv_t m2 = {clampi(vx2, -(1 << 17), (1 << 17) - 1), clampi(vy2, -(1 << 17), (1 << 17) - 1)};
            m2 = to_quarter(m2); /* roundAffinePrecInternal2Amvr(mv, 4): aux_functions.cl:2078-2113 */
            m2.x = shl(m2.x, 2);
            m2.y = shl(m2.y, 2);
            m2 = clip_mv(m2, ctuX + cuXs[c], ctuY + cuYs[c], W, H);
            p.LBx = m2.x;
            p.LBy = m2.y;
        }
        predC[c] = p;
        currC[c] = p;
        bestC[c] = p; /* never read before the first (always successful) update */
        bestCost[c] = (int64_t)1 << 30; /* MAX_LONG = 1<<62 evaluates to 1<<30 in OpenCL C (constants.cl:61) */
        within[c] = (ctuX + cuXs[c] + cuW <= W) && (ctuY + cuYs[c] + cuH <= H);
    }

    /* affine.cl:114-134 / :1066-1098: each item fetches the current samples of its sub-blocks */
    for (int lid = 0; lid < WG; lid++) {
        int c = lid / itemsPerCu;
        for (int pass = 0; pass < nPasses; pass++) {
            int index = pass * stridePerPass + lid % itemsPerCu;
            int sy = (index / sbCols) << 2, sx = (index % sbCols) << 2;
            if (ha && sy >= cuH) break;
            int off = (ctuY + cuYs[c] + sy) * W + ctuX + cuXs[c] + sx;
            if (off < W * H && within[c])
                for (int r = 0; r < 4; r++)
                    for (int q = 0; q < 4; q++) s->cur[lid][pass][r * 4 + q] = cur[off + r * W + q];
        }
    }

    const int numIter = (nCP == 3 ? 4 : 5) + o->extra_grad_iter;
    for (int iter = 0; iter < numIter + 1; iter++) {
        /* ---- prediction + SATD (affine.cl:202-398) ---- */
        for (int lid = 0; lid < WG; lid++) {
            int c = lid / itemsPerCu;
            int64_t acc = 0;
            for (int pass = 0; pass < within[c] * nPasses; pass++) {
                int index = pass * stridePerPass + lid % itemsPerCu;
                int sy = (index / sbCols) << 2, sx = (index % sbCols) << 2;
                if (ha && sy >= cuH) break;
                int isSpread;
                mv_t mv = derive_sub_mv(&currC[c], nCP, cuW, cuH, sx, sy, &isSpread);
                mv = round_mv(mv, 7);                                          /* roundAndClipMv: aux:90-101 */
                mv = clip_mv(mv, ctuX + cuXs[c], ctuY + cuYs[c], W, H);
                int ix = mv.x >> 4, fx = mv.x & 15, iy = mv.y >> 4, fy = mv.y & 15;
                int win[121], pred[16];
                gather_window(ref, W, H, ctuX + cuXs[c] + sx, ctuY + cuYs[c] + sy, ix, iy, win);
                hv_filter(win, fx, fy, pred);
                for (int r = 0; r < 4; r++)
                    for (int q = 0; q < 4; q++)
                        s->tile[(cuYs[c] + sy + r) * CTU + cuXs[c] + sx + q] = (int16_t)pred[r * 4 + q];
                acc += (int64_t)satd4x4(s->cur[lid][pass], pred);
            }
            s->satd[lid] = acc;
        }
        /* ---- per-CU reduction, rate, best update (affine.cl:416-457) ---- */
        for (int c = 0; c < cusPerCtu; c++) {
            int vlid = c * itemsPerCu;
            for (int i = 1; i < itemsPerCu; i++) s->satd[vlid] += s->satd[vlid + i];
            oracle_cpmvs zero;
            memset(&zero, 0, sizeof zero);
            int bits = calc_affine_bits(nCP, &currC[c], nCP == 3 ? &zero : &predC[c]);
            int64_t cost = s->satd[vlid] + (int64_t)get_cost(bits + 2, lambda); /* LOW_DELAY_P: ruiBits = 2 */
            if (cost < bestCost[c]) {
                bestCost[c] = cost;
                bestC[c] = currC[c];
            }
        }
        if (iter == numIter) break;

        /* ---- Sobel over the whole CTU tile (affine.cl:477-494) ---- */
        for (int cs = 0; cs < CTU * CTU; cs++) {
            int x = cs % CTU, y = cs / CTU;
            if (x == 0 || x == CTU - 1 || y == 0 || y == CTU - 1) {
                s->gx[cs] = 0;
                s->gy[cs] = 0;
            } else {
                const int16_t *p = s->tile;
                s->gx[cs] = (int16_t)(p[cs - CTU + 1] - p[cs - CTU - 1] + 2 * p[cs + 1] - 2 * p[cs - 1] + p[cs + CTU + 1] - p[cs + CTU - 1]);
                s->gy[cs] = (int16_t)(p[cs + CTU - 1] - p[cs - CTU - 1] + 2 * p[cs + CTU] - 2 * p[cs - CTU] + p[cs + CTU + 1] - p[cs - CTU + 1]);
            }
        }
        /* ---- CU border replication: rows, then columns, then corners (affine.cl:506-540) ---- */
        for (int c = 0; c < cusPerCtu; c++) {
            int16_t *G[2] = {s->gx, s->gy};
            int o0 = cuYs[c] * CTU + cuXs[c];
            for (int k = 0; k < 2; k++) {
                int16_t *g2 = G[k];
                for (int col = 0; col < cuW; col++) {
                    g2[o0 + col] = g2[o0 + col + CTU];
                    g2[o0 + col + (cuH - 1) * CTU] = g2[o0 + col + (cuH - 2) * CTU];
                }
                for (int row = 0; row < cuH; row++) {
                    g2[o0 + row * CTU] = g2[o0 + row * CTU + 1];
                    g2[o0 + row * CTU + cuW - 1] = g2[o0 + row * CTU + cuW - 2];
                }
                g2[o0] = g2[o0 + CTU + 1];
                g2[o0 + cuW - 1] = g2[o0 + CTU + cuW - 2];
                g2[o0 + (cuH - 1) * CTU] = g2[o0 + (cuH - 2) * CTU + 1];
                g2[o0 + (cuH - 1) * CTU + cuW - 1] = g2[o0 + (cuH - 2) * CTU + cuW - 2];
            }
        }
        /* ---- error tile = current - prediction, in place (affine.cl:547-580) ---- */
        const int ePasses = ha ? (int)ceilf((float)(cuW * cuH * cusPerCtu / 16) / 256) : 4;
        for (int lid = 0; lid < WG; lid++) {
            int c = lid / itemsPerCu;
            for (int pass = 0; pass < ePasses; pass++) {
                int index = pass * stridePerPass + lid % itemsPerCu;
                int sy = (index / sbCols) << 2, sx = (index % sbCols) << 2;
                if (ha && sy >= cuH) break;
                for (int r = 0; r < 4; r++)
                    for (int q = 0; q < 4; q++) {
                        int off = (cuYs[c] + sy + r) * CTU + cuXs[c] + sx + q;
                        s->tile[off] = (int16_t)((int16_t)s->cur[lid][pass][r * 4 + q] - s->tile[off]);
                    }
            }
        }
        /* ---- partial systems per item (affine.cl:671-717) ---- */
        for (int lid = 0; lid < WG; lid++) {
            int c = lid / itemsPerCu;
            int64_t (*pe)[7] = s->eq[lid];
            memset(pe, 0, sizeof(int64_t) * 49);
            for (int pass = 0; pass < (cuW * cuH) / itemsPerCu; pass++) {
                int idx = pass * itemsPerCu + lid % itemsPerCu;
                int j = cuYs[c] + idx / cuW, k = cuXs[c] + idx % cuW;
                int cy = (((idx / cuW) >> 2) << 2) + 2, cx = (((idx % cuW) >> 2) << 2) + 2;
                int gxv = s->gx[j * CTU + k], gyv = s->gy[j * CTU + k];
                int iC[6];
                if (nCP == 3) {
                    iC[0] = gxv; iC[1] = cx * gxv; iC[2] = gyv; iC[3] = cx * gyv; iC[4] = cy * gxv; iC[5] = cy * gyv;
                } else {
                    iC[0] = gxv; iC[1] = cx * gxv + cy * gyv; iC[2] = gyv; iC[3] = cy * gxv - cx * gyv;
                }
                for (int col = 0; col < n; col++) {
                    for (int row = 0; row < n; row++) pe[col + 1][row] += (int64_t)iC[col] * (int64_t)iC[row];
                    pe[col + 1][n] += (int64_t)((uint64_t)((int64_t)iC[col] * (int64_t)s->tile[j * CTU + k]) << 3);
                }
            }
        }
        /* ---- reduce, solve, update (affine.cl:726-893) ---- */
        for (int c = 0; c < cusPerCtu; c++) {
            int vlid = c * itemsPerCu;
            int64_t sum[7][7];
            double dm[7][7];
            memset(dm, 0, sizeof dm);
            for (int col = 1; col < n + 1; col++)
                for (int row = 0; row < 7; row++) sum[col][row] = s->eq[vlid][col][row];
            for (int item = 1; item < itemsPerCu; item++)
                for (int col = 1; col < n + 1; col++)
                    for (int row = 0; row < 7; row++) sum[col][row] += s->eq[vlid + item][col][row];
            for (int col = 1; col < n + 1; col++)
                for (int row = 0; row < 7; row++) dm[col][row] = (double)sum[col][row];
            double a[6];
            oracle_solve(dm, n, o->fused_backsub, a);
            double d[6] = {0, 0, 0, 0, 0, 0};
            d[0] = a[0];
            d[2] = a[2];
            if (nCP == 3) {
                d[1] = a[1] * cuW + a[0];
                d[3] = a[3] * cuW + a[2];
                d[4] = a[4] * cuH + a[0];
                d[5] = a[5] * cuH + a[2];
            } else {
                d[1] = a[1] * cuW + a[0];
                d[3] = -a[3] * cuW + a[2];
            }
            /* scaleDeltaMvs swaps s1/s2 (aux_functions.cl:2203-2210): LT += (d0,d2), RT += (d1,d3), LB += (d4,d5) */
            oracle_cpmvs *cc = &currC[c];
            cc->LTx += scale_delta(d[0], o->cvt_rule);
            cc->LTy += scale_delta(d[2], o->cvt_rule);
            cc->RTx += scale_delta(d[1], o->cvt_rule);
            cc->RTy += scale_delta(d[3], o->cvt_rule);
            cc->LBx += scale_delta(d[4], o->cvt_rule);
            cc->LBy += scale_delta(d[5], o->cvt_rule);
            const int lo = -(1 << 17), hi = (1 << 17) - 1; /* clampCpmvs: aux:2224-2234 */
            cc->LTx = clampi(cc->LTx, lo, hi); cc->LTy = clampi(cc->LTy, lo, hi);
            cc->RTx = clampi(cc->RTx, lo, hi); cc->RTy = clampi(cc->RTy, lo, hi);
            cc->LBx = clampi(cc->LBx, lo, hi); cc->LBy = clampi(cc->LBy, lo, hi);
            mv_t t;                                         /* clipCpmvs: aux:70-86 */
            t.x = cc->LTx; t.y = cc->LTy; t = clip_mv(t, ctuX + cuXs[c], ctuY + cuYs[c], W, H); cc->LTx = t.x; cc->LTy = t.y;
            t.x = cc->RTx; t.y = cc->RTy; t = clip_mv(t, ctuX + cuXs[c], ctuY + cuYs[c], W, H); cc->RTx = t.x; cc->RTy = t.y;
            t.x = cc->LBx; t.y = cc->LBy; t = clip_mv(t, ctuX + cuXs[c], ctuY + cuYs[c], W, H); cc->LBx = t.x; cc->LBy = t.y;
        }
    }
    for (int c = 0; c < cusPerCtu; c++) { /* affine.cl:928-957 */
        gBestCost[base + c] = bestCost[c];
        gBestCpmvs[base + c] = bestC[c];
    }
}

/* ------------------------------------------------------------ public API */

int oracle_num_ctus(int W, int H) { return ((W + 127) / 128) * ((H + 127) / 128); }

void oracle_default_opts(oracle_opts *o) {
    o->extra_grad_iter = 0;
    o->fused_backsub = 1;
    o->cvt_rule = ORACLE_CVT_NVIDIA;
    o->threads = 0;
}

void oracle_ref_pass(const oracle_opts *o, const uint16_t *ref, const uint16_t *cur, int W, int H, float lambda,
                     int64_t *cost[4], oracle_cpmvs *cpmvs[4]) {
    const int nCtus = oracle_num_ctus(W, H);
    for (int pred = 0; pred < 4; pred++) {
        const int ha = pred >= 2, nCP = (pred & 1) ? 3 : 2;
        const int groups = ha ? O_NUM_HA : O_NUM_ALIGNED;
        const oracle_cpmvs *prev = nCP == 3 ? cpmvs[pred - 1] : NULL;
#ifdef _OPENMP
#pragma omp parallel num_threads(o->threads > 0 ? o->threads : omp_get_max_threads())
#endif
        {
            wg_scratch *s = (wg_scratch *)malloc(sizeof(wg_scratch));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
            for (int wg = 0; wg < nCtus * groups; wg++)
                run_wg(o, (const int16_t *)ref, (const int16_t *)cur, W, H, lambda, wg / groups, wg % groups, ha, nCP,
                       cost[pred], cpmvs[pred], prev, s);
            free(s);
        }
    }
}

/* Stage-level entry points for known-answer tests. */
void oracle_predict_4x4(const uint16_t *ref, int W, int H, int px, int py, int mvx, int mvy, int pred[16]) {
    int win[121];
    gather_window((const int16_t *)ref, W, H, px, py, mvx >> 4, mvy >> 4, win);
    hv_filter(win, mvx & 15, mvy & 15, pred);
}
int oracle_satd_4x4(const int org[16], const int pred[16]) { return satd4x4(org, pred); }
int oracle_affine_bits(int nCP, const oracle_cpmvs *c, const oracle_cpmvs *p) { return calc_affine_bits(nCP, c, p); }
int oracle_rate_cost(int bits, float lambda) { return get_cost(bits, lambda); }
int oracle_scale_delta(double d, int cvt_rule) { return scale_delta(d, cvt_rule); }
void oracle_sub_mv(const oracle_cpmvs *c, int nCP, int w, int h, int sx, int sy, int cuX, int cuY, int W, int H, int out[3]) {
    int sp;
    mv_t m = derive_sub_mv(c, nCP, w, h, sx, sy, &sp);
    m = clip_mv(round_mv(m, 7), cuX, cuY, W, H);
    out[0] = m.x; out[1] = m.y; out[2] = sp;
}

/* main_aux_functions.h:1473-1497 */
static int clip3(double mn, double mx, double val) {
    double t = mx < val ? mx : val;
    t = mn > t ? mn : t;
    return (int)floor(t);
}
int oracle_compute_delta_qp(int inputQp, int poc) {
    static const int pocOffset[8] = {1, 5, 4, 5, 4, 5, 4, 5};
    double modelScale = (poc % 8 == 0) ? 0 : 0.259;
    double modelOffset = (poc % 8 == 0) ? 0 : -6.5;
    int qp = inputQp + pocOffset[poc % 8];
    double dQpOffset = qp * modelScale + modelOffset + 0.5;
    qp += clip3(0.0, 3.0, dQpOffset);
    return qp;
}
float oracle_lambda(int inputQp, int poc) { return O_FULL_LAMBDAS[oracle_compute_delta_qp(inputQp, poc)]; }

/* main.cpp:332-335, 584, 591-707: label simulation of the 4-slot reference
 * list.  Call with poc = 1, 2, 3, ... in order on a zero-initialised state
 * (refs = -1).  Returns numRefs; list[] holds the reference POCs, newest first. */
int oracle_ref_list_step(oracle_reflist *st, int poc, int list[4]) {
    int numRefs = poc < 4 ? poc : 4;
    int tempA, tempB;
    int *R = st->refs, *LT = st->is_lt;
    if (poc < 5) {
        tempA = R[0];
        R[0] = poc - 1;
        if (numRefs > 1) { tempB = R[1]; R[1] = tempA; }
        if (numRefs > 2) { tempA = R[2]; R[2] = tempB; }
        if (numRefs > 3) { R[3] = tempA; }
        LT[3] = R[3] % 8 == 0 ? 1 : 0;
    } else {
        int update;
        tempA = R[0];
        R[0] = poc - 1;
        update = LT[1] == 0 ? 1 : (tempA % 8 == 0 && tempA != R[0] ? 1 : 0);
        if (update) {
            tempB = R[1];
            R[1] = tempA;
            update = LT[2] == 0 ? 1 : (tempB % 8 == 0 && tempB != R[1] ? 1 : 0);
            if (update) {
                tempA = R[2];
                R[2] = tempB;
                update = LT[3] == 0 ? 1 : (tempA % 8 == 0 && tempA != R[3] ? 1 : 0);
                if (update) R[3] = tempA;
            }
        }
        LT[3] = R[3] % 8 == 0 ? 1 : 0;
        LT[2] = (R[2] % 8 == 0 && LT[3]) ? 1 : 0;
        LT[1] = (R[1] % 8 == 0 && LT[2]) ? 1 : 0;
    }
    for (int i = 0; i < 4; i++) list[i] = i < numRefs ? R[i] : -1;
    return numRefs;
}
void oracle_ref_list_init(oracle_reflist *st) {
    for (int i = 0; i < 4; i++) { st->refs[i] = -1; st->is_lt[i] = 0; }
}

/* Geometry accessors for table cross-checks: CU k (result index inside a CTU)
 * of prediction class ha -> x, y, w, h. */
int oracle_cu_geometry(int ha, int k, int out[4]) {
    int groups = ha ? O_NUM_HA : O_NUM_ALIGNED;
    for (int g = groups - 1; g >= 0; g--) {
        int st = ha ? O_HA_STRIDE[g] : O_STRIDE[g];
        if (k >= st) {
            int c = k - st;
            int w = ha ? O_HA_W[g] : O_W[g], h = ha ? O_HA_H[g] : O_H[g];
            int ncu = ha ? O_HA_N[g] : (CTU * CTU) / (w * h);
            if (c >= ncu) return -1;
            out[0] = ha ? O_HA_X[g][c] : (c % (CTU / w)) * w;
            out[1] = ha ? O_HA_Y[g][c] : (c / (CTU / w)) * h;
            out[2] = w;
            out[3] = h;
            return g;
        }
    }
    return -1;
}
