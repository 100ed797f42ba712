// qconvex_gpu -- a `qconvex` the UNMODIFIED reference can spawn (wire formats).
//
// convexHullFiles (MAIN:762-806) writes pointList.txt ("3\n<n>\n" then "x y z " lines at 6 significant digits)
// and runs
//     qconvex.exe n  TO "Planes.txt"        < "pointList.txt"        (MAIN:769)
//     qconvex.exe Fv TO "facetVertices.txt" < "pointList.txt"        (MAIN:770)
// then parses Planes.txt (first two lines skipped, 4 reals per facet, MAIN:776-787) and facetVertices.txt
// (first line skipped, "<count> id id id" per facet, first id used, MAIN:791-803).  This program accepts
// those command lines (qhull's `TO file` option; stdout without it), computes the hull on the GPU through the
// C ABI (lqro_convex_hull: facets in qhull's order, apex-first vertex lists) and writes qhull's formats:
//   n : "4\n<nf>\n", then qh_REAL_1 "%6.16g " per coefficient       Fv: "<nf>\n", then "3 a b c"
// Installed under the name `qconvex` (or qconvex.exe) on PATH it turns the reference's file interface into
// a GPU hull without touching its source.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/lqro.h"

int main(int argc, char** argv) {
  bool want_n = false, want_fv = false;
  const char* outp = nullptr;
  for (int a = 1; a < argc; ++a) {
    if (!strcmp(argv[a], "n")) want_n = true;
    else if (!strcmp(argv[a], "Fv")) want_fv = true;
    else if (!strcmp(argv[a], "TO") && a + 1 < argc) outp = argv[++a];
    else { fprintf(stderr, "qconvex_gpu: unsupported option '%s' (supported: n, Fv, TO file)\n", argv[a]); return 2; }
  }
  if (want_n == want_fv) { fprintf(stderr, "usage: qconvex <n|Fv> [TO file] < points\n"); return 2; }
  int dim = 0, n = 0;
  if (scanf("%d %d", &dim, &n) != 2 || dim != 3 || n < 4) { fprintf(stderr, "qconvex_gpu: expected '3 <n>=4..' header\n"); return 1; }
  std::vector<double> pts(3 * (size_t)n);
  for (size_t i = 0; i < pts.size(); ++i)
    if (scanf("%lf", &pts[i]) != 1) { fprintf(stderr, "qconvex_gpu: short input\n"); return 1; }

  lqro_config cfg;
  lqro_default_config(&cfg);
  cfg.n_agents = 2; cfg.max_neighbors = 1;
  cfg.horizon = (n + cfg.n_samples - 1) / cfg.n_samples;  // hull capacity = horizon * n_samples points
  if (cfg.horizon < 1) cfg.horizon = 1;
  cfg.hull_input = LQRO_HULL_FP64;                         // the file already holds the rounded numbers
  lqro_ctx* ctx = nullptr;
  if (lqro_create(&cfg, &ctx) != LQRO_OK) { fprintf(stderr, "qconvex_gpu: %s\n", lqro_last_error(nullptr)); return 3; }
  int nf = 2 * n + 8, degenerate = 0;
  std::vector<double> planes(4 * (size_t)nf);
  std::vector<int> verts(3 * (size_t)nf);
  if (lqro_convex_hull(ctx, pts.data(), n, planes.data(), verts.data(), &nf, &degenerate) != LQRO_OK) {
    fprintf(stderr, "qconvex_gpu: %s\n", lqro_last_error(ctx));
    lqro_destroy(ctx);
    return 4;
  }
  lqro_destroy(ctx);
  if (degenerate) fprintf(stderr, "qconvex_gpu: warning: coplanar/merged facets would differ from qhull (C-0 merging not restated)\n");
  FILE* fo = outp ? fopen(outp, "w") : stdout;
  if (!fo) { perror(outp); return 5; }
  if (want_n) {
    fprintf(fo, "4\n%d\n", nf);
    for (int f = 0; f < nf; ++f) {
      for (int k = 0; k < 4; ++k) fprintf(fo, "%6.16g ", planes[4 * (size_t)f + k]);
      fprintf(fo, "\n");
    }
  } else {
    fprintf(fo, "%d\n", nf);
    for (int f = 0; f < nf; ++f) fprintf(fo, "3 %d %d %d\n", verts[3 * (size_t)f], verts[3 * (size_t)f + 1], verts[3 * (size_t)f + 2]);
  }
  if (outp) fclose(fo);
  return 0;
}
