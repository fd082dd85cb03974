/* TEST INFRASTRUCTURE — definitions for the GLPK / Armadillo stubs: every one
 * aborts.  See glpk.h / armadillo in this directory. */
#include <cstdio>
#include <cstdlib>
#include "glpk.h"
#include "armadillo"
#define DIE() do { std::fprintf(stderr, "oracle/_ref: %s is a stub (training-side dependency not installed)\n", __func__); std::abort(); } while (0)
extern "C" {
glp_prob *glp_create_prob(void) { DIE(); }
void glp_delete_prob(glp_prob *) { DIE(); }
void glp_init_iocp(glp_iocp *) { DIE(); }
void glp_init_smcp(glp_smcp *) { DIE(); }
void glp_set_obj_dir(glp_prob *, int) { DIE(); }
int glp_add_rows(glp_prob *, int) { DIE(); }
int glp_add_cols(glp_prob *, int) { DIE(); }
void glp_set_obj_coef(glp_prob *, int, double) { DIE(); }
void glp_set_col_kind(glp_prob *, int, int) { DIE(); }
void glp_set_col_bnds(glp_prob *, int, int, double, double) { DIE(); }
void glp_set_row_bnds(glp_prob *, int, int, double, double) { DIE(); }
void glp_load_matrix(glp_prob *, int, const int *, const int *, const double *) { DIE(); }
int glp_get_row_type(glp_prob *, int) { DIE(); }
double glp_get_row_lb(glp_prob *, int) { DIE(); }
double glp_get_row_ub(glp_prob *, int) { DIE(); }
int glp_intopt(glp_prob *, const glp_iocp *) { DIE(); }
int glp_simplex(glp_prob *, const glp_smcp *) { DIE(); }
int glp_exact(glp_prob *, const glp_smcp *) { DIE(); }
double glp_mip_obj_val(glp_prob *) { DIE(); }
double glp_mip_col_val(glp_prob *, int) { DIE(); }
double glp_get_obj_val(glp_prob *) { DIE(); }
double glp_get_col_prim(glp_prob *, int) { DIE(); }
int glp_term_out(int) { DIE(); }
}
namespace arma {
bool kmeans(fmat &, const fmat &, int, seed_mode, int, bool) { DIE(); }
}
