/* include/ldpc_toolbox.h — C ABI of the B200-native LDPC decoder / BER engine.
 *
 * Part 1 re-declares, with identical names, argument order and types, the nine entry points of
 * the reference's FFI (reference include/ldpc_toolbox.h:11-30, implemented in
 * src/c_api/decoder.rs:78-137 and src/c_api/encoder.rs:54-97), so a program linked against the
 * reference's libldpc_toolbox.{so,a} links against this library unchanged.
 *
 * Part 2 is additive: batched decode (the native shape of the GPU path), device-pointer
 * variants, an on-device BER engine and introspection.  Nothing in part 1 changes meaning.
 *
 * Error behaviour: constructors return NULL on any error, as the reference does
 * (c_api/decoder.rs:84-87).  Where the reference panics across the FFI (length mismatches,
 * flooding.rs:56, c_api/decoder.rs:51,:61, c_api/encoder.rs:44,:48) this library returns -2
 * from decode / leaves the encoder output untouched, and never reads or writes out of bounds.
 * ldpc_toolbox_last_error() describes the last failure on the calling thread.
 *
 * There is no CPU fallback: without a CUDA device every constructor returns NULL.
 */
#ifndef LDPC_TOOLBOX_H_
#define LDPC_TOOLBOX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Part 1 — the reference's ABI
 * ------------------------------------------------------------------------------------------ */

/* reference src/c_api/decoder.rs:78-88.  `implementation` is one of the 36 names of
 * src/decoder/factory.rs:240-277 (case-sensitive); `puncturing` is "" or e.g. "1,1,1,1,0". */
void *ldpc_toolbox_decoder_ctor(const char *alist_file_path, const char *implementation,
                                const char *puncturing);
/* reference src/c_api/decoder.rs:90-102 */
void *ldpc_toolbox_decoder_ctor_alist_string(const char *alist, const char *implementation,
                                             const char *puncturing);
/* reference src/c_api/decoder.rs:104-107 */
void ldpc_toolbox_decoder_dtor(void *decoder);
/* reference src/c_api/decoder.rs:109-122.  Returns the iteration count (>= 0) on success and -1
 * on decoding failure; `output` receives the first output_len hard bits (one 0/1 per byte) in
 * both cases.  -2: argument error / a check node the min* rules cannot process. */
int32_t ldpc_toolbox_decoder_decode_f64(void *decoder, uint8_t *output, size_t output_len,
                                        const double *llrs, size_t llrs_len,
                                        uint32_t max_iterations);
/* reference src/c_api/decoder.rs:124-137 (f32 is widened to f64, :69-72) */
int32_t ldpc_toolbox_decoder_decode_f32(void *decoder, uint8_t *output, size_t output_len,
                                        const float *llrs, size_t llrs_len,
                                        uint32_t max_iterations);

/* reference src/c_api/encoder.rs:54-65 */
void *ldpc_toolbox_encoder_ctor(const char *alist_file_path, const char *puncturing);
/* reference src/c_api/encoder.rs:67-77 */
void *ldpc_toolbox_encoder_ctor_alist_string(const char *alist, const char *puncturing);
/* reference src/c_api/encoder.rs:79-82 */
void ldpc_toolbox_encoder_dtor(void *encoder);
/* reference src/c_api/encoder.rs:84-97.  input bytes equal to 1 are ones, anything else zero;
 * output_len must equal the (punctured) codeword length. */
void ldpc_toolbox_encoder_encode(void *encoder, uint8_t *output, size_t output_len,
                                 const uint8_t *input, size_t input_len);

/* ------------------------------------------------------------------------------------------
 * Part 2 — additive entry points
 * ------------------------------------------------------------------------------------------ */

const char *ldpc_toolbox_last_error(void);

/* Constructor with placement options: CUDA device ordinal (-1 = current) and the maximum number
 * of 128-frame tiles processed per kernel launch (0 = automatic). */
void *ldpc_toolbox_decoder_ctor_ex(const char *alist, int alist_is_path, const char *implementation,
                                   const char *puncturing, int device, int max_tiles);

/* Batched decode(llrs, max_iterations) on host buffers: frame f reads llrs[f*llrs_len ..] and
 * writes output[f*output_stride .. +output_len] and iterations[f] (>= 0, -1 failure, -2 error).
 * Returns 0, or -2 on an argument / CUDA error (nothing is written). */
int32_t ldpc_toolbox_decoder_decode_batch_f32(void *decoder, uint8_t *output, size_t output_len,
                                              size_t output_stride, const float *llrs,
                                              size_t llrs_len, size_t nframes,
                                              uint32_t max_iterations, int32_t *iterations);
int32_t ldpc_toolbox_decoder_decode_batch_f64(void *decoder, uint8_t *output, size_t output_len,
                                              size_t output_stride, const double *llrs,
                                              size_t llrs_len, size_t nframes,
                                              uint32_t max_iterations, int32_t *iterations);
/* Asynchronous form of the batched decode, for streaming callers: submit returns a ticket (>= 0) as soon
 * as the copies and kernels are enqueued (-2 on an argument / CUDA error) and wait(ticket) blocks until
 * that batch — and every batch submitted before it — has been written to `output` / `iterations`.  The
 * library pipelines H2D copies, decoding and D2H copies across consecutive submits, so a caller that keeps
 * two batches in flight never exposes a copy.  All buffers of a submit must stay valid until its wait
 * returns, and must be pinned (cudaHostAlloc / cudaHostRegister) for the call not to block.
 * decode_batch_* is submit + wait. */
int64_t ldpc_toolbox_decoder_submit_batch_f32(void *decoder, uint8_t *output, size_t output_len,
                                              size_t output_stride, const float *llrs,
                                              size_t llrs_len, size_t nframes,
                                              uint32_t max_iterations, int32_t *iterations);
int64_t ldpc_toolbox_decoder_submit_batch_f64(void *decoder, uint8_t *output, size_t output_len,
                                              size_t output_stride, const double *llrs,
                                              size_t llrs_len, size_t nframes,
                                              uint32_t max_iterations, int32_t *iterations);
int32_t ldpc_toolbox_decoder_wait(void *decoder, int64_t ticket);

/* One handle = one caller: like the reference's `&mut` handle (src/c_api/decoder.rs:120) a decoder must not be
 * used from two threads at once, and the device-pointer calls below share one workspace per handle — calls
 * on the same handle must be issued on ONE stream (use one handle per stream).  Every entry point restores
 * the caller's current CUDA device before returning.
 *
 * Same, with every buffer in device memory of the decoder's GPU; asynchronous on `cuda_stream`
 * (a cudaStream_t; NULL = the legacy default stream). */
int32_t ldpc_toolbox_decoder_decode_batch_device_f32(void *decoder, uint8_t *d_output,
                                                     size_t output_len, size_t output_stride,
                                                     const float *d_llrs, size_t llrs_len,
                                                     size_t nframes, uint32_t max_iterations,
                                                     int32_t *d_iterations, void *cuda_stream);
int32_t ldpc_toolbox_decoder_decode_batch_device_f64(void *decoder, uint8_t *d_output,
                                                     size_t output_len, size_t output_stride,
                                                     const double *d_llrs, size_t llrs_len,
                                                     size_t nframes, uint32_t max_iterations,
                                                     int32_t *d_iterations, void *cuda_stream);

/* Test hook for the float decoders (flooding and frame-per-CTA layered kernels): batched decode that also returns
 * every frame's posterior LLRs as f64 [nframes][n] — the flooding decoder's output_llrs (reference
 * src/decoder/flooding.rs:111-125) / the layered decoder's Qv (src/decoder/horizontal_layered.rs:65-88) when the
 * frame stopped.  Frames with 0 iterations have none (zeros).  One chunk only; -2 on error. */
int32_t ldpc_toolbox_decoder_decode_batch_posteriors_f32(void *decoder, uint8_t *output, size_t output_len,
                                                         size_t output_stride, const float *llrs, size_t llrs_len,
                                                         size_t nframes, uint32_t max_iterations,
                                                         int32_t *iterations, double *posteriors);
int32_t ldpc_toolbox_decoder_decode_batch_posteriors_f64(void *decoder, uint8_t *output, size_t output_len,
                                                         size_t output_stride, const double *llrs, size_t llrs_len,
                                                         size_t nframes, uint32_t max_iterations,
                                                         int32_t *iterations, double *posteriors);

/* Introspection */
size_t ldpc_toolbox_decoder_codeword_len(void *decoder);   /* n (columns of H) */
size_t ldpc_toolbox_decoder_info_len(void *decoder);       /* k = n - rows of H */
size_t ldpc_toolbox_decoder_num_edges(void *decoder);      /* ones in H */
size_t ldpc_toolbox_decoder_llrs_len(void *decoder);       /* expected llrs_len (after puncturing) */
/* device time in ms of the stages of the last processed chunk: [ingest, decode, emit];
 * returns the number of kernels launched by this handle so far */
int64_t ldpc_toolbox_decoder_last_timing(void *decoder, float *ms3);
/* average device time in ms of the BP kernel over the launches (flooding / layered-tile kernels) since the previous
 * call — at most the last 32 — and their number; CUDA events on the launching stream, resolved here, so a caller
 * can time back-to-back calls without synchronising after each of them */
float ldpc_toolbox_decoder_average_decode_ms(void *decoder, int64_t *launches);

/* On-device BER Monte-Carlo engine (BPSK/AWGN), the GPU counterpart of the reference's
 * BerTest/Worker loop (reference src/simulation/ber.rs:246-282, :297-368, :436-481).
 * ldpc_toolbox_ber_run simulates the global frames [first_frame, first_frame+nframes) at one
 * Eb/N0 and ADDS to counters[9] = {frames, bit_errors, frame_errors, false_decodes,
 * total_iterations, correct_iterations, bch_bit_errors, bch_frame_errors,
 * bch_correct_iterations}.  Frames are keyed by their global index (Philox counter), so any
 * sharding of the index range over calls or GPUs gives the same totals.  Returns 0 or -2. */
void *ldpc_toolbox_ber_ctor(const char *alist, int alist_is_path, const char *implementation,
                            const char *puncturing, int device, int max_tiles);
void ldpc_toolbox_ber_dtor(void *ber);
/* Modulation of the simulated link: "BPSK" (default) or "8PSK" (reference src/simulation/factory.rs:56-86,
 * src/simulation/modulation.rs:144-288: DVB-S2 Gray mapping, complex AWGN, exact max* demapper), and the
 * DVB-S2 bit interleaver (reference src/simulation/interleaving.rs:40-85, src/simulation/ber.rs:250-252):
 * interleaving_columns = 0 none, n > 0 n columns, n < 0 |n| columns with rows read backwards.
 * Returns 0, or -2 (unknown name; frame length not a multiple of 3 bits / of the columns — the
 * reference panics on those at the first frame). */
int32_t ldpc_toolbox_ber_set_modulation(void *ber, const char *modulation, int32_t interleaving_columns);
int32_t ldpc_toolbox_ber_run(void *ber, float ebn0_db, uint32_t max_iterations,
                             uint64_t first_frame, uint64_t nframes, uint64_t seed,
                             uint64_t bch_max_errors, uint64_t *counters);
/* Asynchronous form: submit enqueues one batch (front-end, decode, back-end, counter read-back) and returns a
 * ticket (>= 0; -2 on error) without waiting; wait(ticket) blocks until that batch is done and ADDS its nine
 * counters.  At most two tickets may be in flight per engine (two lanes with their own stream, buffers and
 * decoder workspace), so the caller's stop rule and counter reduction for batch i overlap the kernels of batch
 * i+1 — the GPU counterpart of the reference's controller thread receiving results while its workers run ahead
 * (reference src/simulation/ber.rs:312-343).  ldpc_toolbox_ber_run = submit + wait. */
int64_t ldpc_toolbox_ber_submit(void *ber, float ebn0_db, uint32_t max_iterations, uint64_t first_frame,
                                uint64_t nframes, uint64_t seed, uint64_t bch_max_errors);
int32_t ldpc_toolbox_ber_wait(void *ber, int64_t ticket, uint64_t *counters);
/* test hook: same as ldpc_toolbox_ber_run, also copying out (any pointer may be NULL) the f32
 * LLRs [nframes][n_tx], decoded info bytes [nframes][k], iterations and packed messages */
int32_t ldpc_toolbox_ber_run_dump(void *ber, float ebn0_db, uint32_t max_iterations,
                                  uint64_t first_frame, uint64_t nframes, uint64_t seed,
                                  uint64_t bch_max_errors, uint64_t *counters, float *llrs,
                                  uint8_t *decoded, int32_t *iterations, uint32_t *messages);
/* what[0]=k, [1]=N_cw, [2]=N (transmitted symbols per frame) */
void ldpc_toolbox_ber_dims(void *ber, uint64_t *what3);
double ldpc_toolbox_ber_rate(void *ber);                       /* k / N, ber.rs:259 */
double ldpc_toolbox_ber_noise_sigma(void *ber, float ebn0_db); /* ber.rs:300-302 */

int32_t ldpc_toolbox_num_implementations(void);
const char *ldpc_toolbox_implementation_name(int32_t index);

#ifdef __cplusplus
}
#endif

#endif /* LDPC_TOOLBOX_H_ */
