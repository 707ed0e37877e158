// ldpc_toolbox_b200/csrc/flood_i8_w16.cu — K1 with one 16-line stage per warp (codes with check degrees up to 16,
// e.g. DVB-S2 rates 3/5, 2/3, 3/4): a separate translation unit so the three stage capacities build in parallel.
#define LDPC_I8_WCAP 16
#include "flood_i8.cu"
