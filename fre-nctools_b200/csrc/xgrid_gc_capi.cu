// Great-circle exchange-grid path (create_xgrid_great_circle, create_xgrid.c:1366-1466).
#include "xgrid_plan.h"
long long xgb_generate_great_circle(xgb_plan* p, int order)
{
  (void)p; (void)order;
  xgb_set_error("great-circle algorithm: not built into this library yet");
  return -1;
}
