/* placeholder replaced below */
#include "xgrid_oracle.h"
long orc_create_xgrid_great_circle(int a, int b, int c, int d, const double *e, const double *f, const double *g, const double *h, const double *m, long cap, int *i1, int *j1, int *i2, int *j2, double *xa, double *xc, double *yc) { return -1; }
