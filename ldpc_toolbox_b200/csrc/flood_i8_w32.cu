// ldpc_toolbox_b200/csrc/flood_i8_w32.cu — K1 with one 32-line stage per warp (check degrees up to 32: DVB-S2 rates
// 4/5 ... 9/10): a separate translation unit so the three stage capacities build in parallel.
#define LDPC_I8_WCAP 32
#include "flood_i8.cu"
