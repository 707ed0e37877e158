// ldpc_toolbox_b200/csrc/decoder_impl.hpp — internal launch interfaces between the host-side
// decoder object (decoder.cu) and the kernels (flood_i8.cu, flood_float.cu, layered.cu, ingest.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "device_common.cuh"

namespace ldpc {

void set_last_error(const std::string& msg);
const std::string& last_error();

// ---- ingest.cu ---------------------------------------------------------------------------------
struct IngestLaunch {
    const void* llrs;        // device, [nframes][llrs_len] f32 or f64 (frame-major, as the caller gave them)
    bool is_f64;
    size_t llrs_len;         // per frame (punctured length when a puncturer is attached)
    size_t nframes;          // frames present; tiles are padded with LLR=+1 (clean all-zero word)
    int n;                   // codeword length
    const int* src_map;      // device, n entries: index into the frame's llrs or -1 (punctured -> 0.0); may be null
    int num_tiles;
    int words_per_lane;      // 1: 128-frame tiles, 4: 512-frame tiles (int8 decoders only)
    // outputs (any may be null)
    uint32_t* inq_i8;        // [tiles][n][32][NW] int8x4: quantised LLRs (arithmetic.rs:690-699)
    float* in_f32;           // [tiles][n][128] f32 (`llr as f32`)
    double* in_f64;          // [tiles][n][128] f64
    int16_t* in_i16;         // [tiles][n][128] quantised LLRs widened to i16 (layered VarLlr, arithmetic.rs:709-711)
    void* hard;              // [tiles][n][32] raw-sign hard decisions (x <= 0.0), 4*NW bits per lane (u8 / u16)
};
bool launch_ingest(const IngestLaunch& L, cudaStream_t stream);

struct EmitLaunch {
    const void* final_hard;      // [tiles][n][32] u8 (NW=1) or u16 (NW=4)
    int n;
    int num_tiles;
    int words_per_lane;
    size_t nframes;
    uint8_t* out;                // device, [nframes][out_stride] one 0/1 byte per bit
    size_t out_len, out_stride;
};
bool launch_emit(const EmitLaunch& L, cudaStream_t stream);
// test hook: posteriors of the float flooding kernel, tiles [tiles][n][128] (f32 / f64) -> [nframes][n] f64
bool launch_emit_posteriors(const void* post_tiles, bool is_f64, int n, size_t nframes, double* out, cudaStream_t stream);

// Variables grouped by degree (host-built, device-resident): class k holds the variables
// var_list[off[k] .. off[k+1]) which all have degree deg[k] (1..8), with their row-major edge ids
// flattened in cols[v] order at var_edges[edge_off[k] + i*deg[k] + j].  One trailing class with
// deg[k] = 0 collects every variable of degree > 8 (edges looked up through col_ptr / col_edge).
// Degree-0 variables are in no class.
struct VarClasses {
    int num_classes;
    int deg[10];
    int off[11];
    int edge_off[10];
    const int* var_list;
    const int* var_edges;
};

// ---- flood_i8.cu -------------------------------------------------------------------------------
// Staircase fusion (DVB-S2 and other IRA codes, reference src/codes/dvbs2.rs:92-96): a degree-2 variable whose
// two checks are consecutive rows r, r+1 — last slot of row r, second-to-last slot of row r+1 — is updated
// inside the check pass, right after both of its incoming messages of the iteration exist, by the warp that
// owns both rows; its messages then cross HBM twice per iteration instead of four times.  Exact under
// flooding: the variable update of iteration i depends only on the two check outputs of iteration i.
// Rows are dealt to warps in chunks of chunk_rows consecutive rows; a staircase variable that straddles
// two chunks stays in the ordinary variable pass.
constexpr int kFuseChunkRowsDefault = 32;
constexpr int kFuseMaxRowDeg = 8;
struct RowMeta {             // one 16-byte record per check row
    int e0;                  // first row-major edge
    int d_flags;             // bits 0-15 degree, bit 16: slot d-2 is fused with row r-1, bit 17: slot d-1 is fused with row r+1
    int fuse_var;            // variable of the slot fused with row r-1, or -1 (always r + fuse_var_off: the kernel uses that)
    int last_var;            // variable of the last slot (raw-sign line of the fused variable at start-up)
};
struct FloodI8Launch {
    DeviceGraph graph;
    VarClasses classes;
    const RowMeta* row_meta; // m records
    const int* snap_src;     // n entries: >= 0 first row-major edge of the variable, <= -2 fused (row -2 - x), -1 no checks
    int snap_n;              // variables whose final hard decisions are read back (output_len)
    void* cbit;              // [tiles][2][m][32] hard decisions of the fused variables, by iteration parity (u8 / u16 per lane)
    int chunk_rows;          // rows per chunk (power of two), the same value the row records were built with
    int fuse_var_off;        // every fused variable satisfies v = (row of its second check) + fuse_var_off
    int graph_max_row_deg;   // largest check degree (rows beyond the stage capacity keep an explicit initialisation)
    int num_tiles;
    int words_per_lane;      // 1 or 4
    uint32_t* msg;
    void* hbit;
    const uint32_t* inq;
    const void* raw0;
    void* final_hard;
    int32_t* iters;
    int max_iter;
    bool aminstar, jones, hardlimit, deg1clip;
    int cluster;             // CTAs per tile (thread-block cluster of 1 .. 16): small batches fill the GPU this way
    int wide_cap;            // 0: every row fits the register path; 16 / 32: kernel with one 16- / 32-line stage per warp
};
bool launch_flood_i8(const FloodI8Launch& L, cudaStream_t stream);
int flood_i8_max_row_degree();

// ---- generic_bp.cu -----------------------------------------------------------------------------
struct GenericLaunch {
    DeviceGraph graph;
    int num_tiles;
    int rule;               // RuleId of rules.cuh
    bool is_f64, is_i8, hardlimit;
    void* msg;              // flooding: messages F [tiles][E][128]; layered: Rcv (F or int8) [tiles][E][128]
    uint8_t* hbit;          // flooding only: [tiles][E][32]
    const void* in;         // flooding: channel LLRs F [tiles][n][128]
    void* in_out_q;         // layered: Qv (F or int16) [tiles][n][128], initialised by the ingest kernel
    const uint8_t* raw0;    // [tiles][n][32]
    uint8_t* final_hard;    // [tiles][n][32]
    int32_t* iters;
    int max_iter;
    const int* level_ptr;   // layered: level schedule
    const int* level_rows;
    int num_levels;
    int cluster;            // flooding: CTAs per tile (thread-block cluster of 1, 2, 4 or 8), small batches fill the GPU this way
    void* post;             // flooding, test hook (may be null): posteriors F [tiles][n][128] of each frame's last processed iteration
};
bool launch_flood_float(const GenericLaunch& L, cudaStream_t stream);
bool launch_layered(const GenericLaunch& L, cudaStream_t stream);
int generic_max_row_degree();

// ---- layered_smem.cu ---------------------------------------------------------------------------
// Level schedule in ELL order (host-built once per code): rows are listed level by level; the slots
// of a level form a block [max degree of the level][rows of the level], so slot j of consecutive
// rows of a level is contiguous.  Rcv of a frame uses the same indexing.
#ifndef LDPC_K3Q_MAX_THREADS
#define LDPC_K3Q_MAX_THREADS 384
#endif
// CTA size bound of K3q.  384 threads x 2 CTAs per SM (80 registers, a few spills) beat 256 x 2 at 107
// registers by 15-25 % on 5G-NR Z=384: the kernel is latency-bound, and 384 threads take exactly one row
// each of a 384-row level (256 leave half the CTA idle in the second round).
constexpr int kSmemLayeredMaxThreads = LDPC_K3Q_MAX_THREADS;

struct LayeredSmemGraph {
    int n, m, num_levels;
    const int* level_ptr;    // num_levels+1, into the level-ordered row arrays below
    const int* level_ell;    // num_levels: ELL index of the level's block
    const int* level_deg;    // num_levels: the common degree of the level's rows, or -1 if they differ
    const int* row_deg;      // m (level order): degree of each row; slot j of row i of level l is at level_ell[l] + j*rows(l) + i
    const int* ell_col;      // ell_size: variable of each slot (padding slots are never read)
    size_t ell_size;
};
struct LayeredSmemLaunch {
    LayeredSmemGraph graph;
    int rule;
    bool is_f64, is_i8, hardlimit;
    int threads;             // CTA size (multiple of 32)
    const void* llrs;        // device, caller layout [nframes][llrs_len]
    bool in_f64;
    size_t llrs_len, nframes;
    const int* src_map;      // depuncture map or null
    void* rcv;               // [nframes][ell_size] check->variable messages (f32 / f64 / int8)
    uint8_t* out;            // device, [nframes][out_stride]
    size_t out_len, out_stride;
    int32_t* iters;          // device, [nframes]
    int max_iter;
    double* post;            // test hook (may be null): device, [nframes][n] posteriors (Qv) at the end of each frame's decode
};
bool launch_layered_smem(const LayeredSmemLaunch& L, cudaStream_t stream);
size_t layered_smem_bytes(int n, bool is_f64, bool is_i8);
size_t layered_smem_rcv_elem(bool is_f64, bool is_i8);

}  // namespace ldpc
