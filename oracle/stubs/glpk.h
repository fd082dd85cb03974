/* TEST INFRASTRUCTURE — stub of the GLPK C API, used only to compile the
 * UNMODIFIED reference sources into oracle/_ref (GLPK is not installed here).
 * GLPK is training-side only (reference VAQ.cpp:339-524, BitVecEngine.hpp:339-507,
 * 640-809); no function on the query-time hot path calls it.  Every entry point
 * aborts (oracle/stubs/stubdefs.cpp) so an accidental call is loud. */
#ifndef VAQ_B200_STUB_GLPK_H
#define VAQ_B200_STUB_GLPK_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct glp_prob glp_prob;
typedef struct { int msg_lev; int presolve; int tm_lim; int pad[32]; } glp_iocp;
typedef struct { int msg_lev; int meth; int presolve; int tm_lim; int pad[32]; } glp_smcp;
#define GLP_ON 1
#define GLP_OFF 0
#define GLP_MIN 1
#define GLP_MAX 2
#define GLP_CV 1
#define GLP_IV 2
#define GLP_BV 3
#define GLP_FR 1
#define GLP_LO 2
#define GLP_UP 3
#define GLP_DB 4
#define GLP_FX 5
#define GLP_PRIMAL 1
#define GLP_DUALP 2
#define GLP_DUAL 3
#define GLP_PT_STD 0x11
#define GLP_RT_STD 0x11
glp_prob *glp_create_prob(void);
void glp_delete_prob(glp_prob *);
void glp_init_iocp(glp_iocp *);
void glp_init_smcp(glp_smcp *);
void glp_set_obj_dir(glp_prob *, int);
int glp_add_rows(glp_prob *, int);
int glp_add_cols(glp_prob *, int);
void glp_set_obj_coef(glp_prob *, int, double);
void glp_set_col_kind(glp_prob *, int, int);
void glp_set_col_bnds(glp_prob *, int, int, double, double);
void glp_set_row_bnds(glp_prob *, int, int, double, double);
void glp_load_matrix(glp_prob *, int, const int *, const int *, const double *);
int glp_get_row_type(glp_prob *, int);
double glp_get_row_lb(glp_prob *, int);
double glp_get_row_ub(glp_prob *, int);
int glp_intopt(glp_prob *, const glp_iocp *);
int glp_simplex(glp_prob *, const glp_smcp *);
int glp_exact(glp_prob *, const glp_smcp *);
double glp_mip_obj_val(glp_prob *);
double glp_mip_col_val(glp_prob *, int);
double glp_get_obj_val(glp_prob *);
double glp_get_col_prim(glp_prob *, int);
int glp_term_out(int);
#ifdef __cplusplus
}
#endif
#endif
