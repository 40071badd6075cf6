/* placeholder; filled in below */
#include <stdint.h>
struct orc_params;
int orc_sgbm_compute(const uint8_t *l, int ls, const uint8_t *r, int rs, int W, int H,
                     const struct orc_params *p, int16_t *disp, int dstep) { return -38; }
