// Conservative apply path (do_scalar_conserve_interp, conserve_interp.c:507-910).
#include "xgrid_plan.h"
void xgb_apply_release(xgb_plan* p) { (void)p; }
