/*
 * ame_oracle.h -- interface of the CPU parity oracle (TEST INFRASTRUCTURE ONLY;
 * see the header of ame_oracle.c).  Plain C, loaded by tests through ctypes.
 */
#ifndef AME_ORACLE_H
#define AME_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* == Cpmvs of /root/reference/typedef.h:6-9 (28 bytes; nCPs is never written) */
typedef struct { int32_t nCPs, LTx, LTy, RTx, RTy, LBx, LBy; } oracle_cpmvs;

enum { ORACLE_CVT_X86 = 0, ORACLE_CVT_NVIDIA = 1 };

typedef struct {
    int extra_grad_iter; /* --ExtraGradientIter */
    int fused_backsub;   /* 1: back-substitution accumulates with fma (OpenCL FP_CONTRACT ON) */
    int cvt_rule;        /* double->int rule for degenerate systems */
    int threads;         /* OpenMP threads, 0 = all */
} oracle_opts;

typedef struct { int refs[4]; int is_lt[4]; } oracle_reflist;

void oracle_default_opts(oracle_opts *o);
int oracle_num_ctus(int W, int H);
/* One reference pass = FULL_2CP, FULL_3CP, HALF_2CP, HALF_3CP (index 0..3).
 * cost[p]/cpmvs[p] have nCtus*201 (p<2) or nCtus*284 (p>=2) entries. */
void oracle_ref_pass(const oracle_opts *o, const uint16_t *ref, const uint16_t *cur, int W, int H, float lambda,
                     int64_t *cost[4], oracle_cpmvs *cpmvs[4]);

void oracle_solve(double m[7][7], int n, int fused_backsub, double out[6]);
void oracle_predict_4x4(const uint16_t *ref, int W, int H, int px, int py, int mvx, int mvy, int pred[16]);
int oracle_satd_4x4(const int org[16], const int pred[16]);
int oracle_affine_bits(int nCP, const oracle_cpmvs *c, const oracle_cpmvs *p);
int oracle_rate_cost(int bits, float lambda);
int oracle_scale_delta(double d, int cvt_rule);
void oracle_sub_mv(const oracle_cpmvs *c, int nCP, int w, int h, int sx, int sy, int cuX, int cuY, int W, int H, int out[3]);
int oracle_compute_delta_qp(int inputQp, int poc);
float oracle_lambda(int inputQp, int poc);
void oracle_ref_list_init(oracle_reflist *st);
int oracle_ref_list_step(oracle_reflist *st, int poc, int list[4]);
int oracle_cu_geometry(int ha, int k, int out[4]);

#ifdef __cplusplus
}
#endif
#endif
