// C wrappers around the UNMODIFIED reference header scan2mesh/mesh_distance/sample2meshdist.h (compiled where it lies
// under /root/reference, against oracle/eigen_shim).  Test infrastructure: pins oracle/mesh_distance.py.
#include "sample2meshdist.h"

extern "C" {
// kind: 0 distance, 1 squared distance, 2 Geman-McClure of the squared distance (sigma); part: 0 plane, 1..3 edges
// ab / bc / ca, 4..6 vertices a / b / c (sample2meshdist.h:182-195).  The derivative buffers are accumulated into.
double s2m_tri(int kind, double sigma, int part, const double *x, const double *a, const double *b, const double *c,
               double *dx, double *da, double *db, double *dc) {
    if (kind == 0) { instances::Distance d; return d.tri(part, x, a, b, c, dx, da, db, dc); }
    if (kind == 1) { instances::SquaredDistance d; return d.tri(part, x, a, b, c, dx, da, db, dc); }
    instances::GMDistance d(sigma);
    return d.tri(part, x, a, b, c, dx, da, db, dc);
}
}
