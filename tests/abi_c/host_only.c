/* TEST: the header is plain C (compiled as strict C99) and the library links from a C program.  Calls only the
 * host-side entry points (no GPU needed): defaults, mesh size, field count, table capacity, and - without a
 * device - that a compute context is refused rather than emulated. */
#include <stdio.h>

#include "eigensolver_b200.h"

int main(void) {
    esb_model m;
    int32_t n = 0, nf = 0, mx = 0;
    esb_context* ctx = NULL;
    int rc;
    if (esb_model_defaults(ESB_CYLINDER_DENSITY, &m)) return 1;
    if (esb_mesh_size(&m, &n) || esb_model_n_fields(&m, &nf) || esb_model_max_steps(&m, &mx)) return 2;
    if (esb_sizeof_model() != (int)sizeof(esb_model)) return 3;
    rc = esb_create(0, &ctx);
    printf("%d %d %d %d %d %d\n", esb_version(), (int)m.n_steps, (int)n, (int)nf, (int)mx, rc);
    if (rc == ESB_OK) esb_destroy(ctx);
    return 0;
}
