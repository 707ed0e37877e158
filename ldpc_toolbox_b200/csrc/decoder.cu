// ldpc_toolbox_b200/csrc/decoder.cu — host-side decoder object: owns the uploaded graph layout and
// the per-tile HBM workspace, and drives ingest -> BP kernel -> emit on one CUDA stream.
//
// Mirrors DecoderImplementation::build_decoder (reference src/decoder/factory.rs:202-208) and the
// decode wrapper of the C API (reference src/c_api/decoder.rs:50-72).
#include "decoder.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>

#include "decoder_impl.hpp"
#include "rules.cuh"

namespace ldpc {

namespace {
thread_local std::string g_last_error;
}
void set_last_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error() { return g_last_error; }

namespace {

// makes the decoder's GPU current for a call and restores the caller's device afterwards
struct DeviceScope {
    int prev = -1;
    bool ok = true;
    explicit DeviceScope(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t count = 0;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; count = 0; }
    bool ensure(size_t c) {
        if (c <= count) return true;
        release();
        if (cudaMalloc(&p, c * sizeof(T)) != cudaSuccess) {
            cudaGetLastError();
            set_last_error("cudaMalloc of " + std::to_string(c * sizeof(T)) + " bytes failed");
            return false;
        }
        count = c;
        return true;
    }
    bool upload(const std::vector<T>& h) {
        if (!ensure(std::max<size_t>(h.size(), 1))) return false;
        if (!h.empty() && cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
            set_last_error("graph upload failed");
            return false;
        }
        return true;
    }
};

// per-launch decoder state in HBM; the host path keeps two of them so that two
// launches can be resident at once
struct Workspace {
    DevBuf<uint8_t> msg, inq, hbit, hard, final_hard, cbit;
    DevBuf<int32_t> iters_tile;
    size_t bytes() const { return msg.count + inq.count + hbit.count + hard.count + final_hard.count + cbit.count; }
};

// one staging slot of the host path: H2D target, D2H source and the events that order their reuse
struct StageSlot {
    DevBuf<uint8_t> in, out;
    DevBuf<int32_t> iters;
    cudaEvent_t copied = nullptr, ingested = nullptr, emitted = nullptr, drained = nullptr;
    bool used = false;
};

class GpuDecoder final : public LdpcDecoder {
public:
    GpuDecoder(const DecoderImplementation& impl, const Graph& g) : impl_(impl), g_(g) {}
    ~GpuDecoder() override {
        cudaSetDevice(device_);
        cudaDeviceSynchronize();
        if (stream_) cudaStreamDestroy(stream_);
        if (stream2_) cudaStreamDestroy(stream2_);
        if (h2d_stream_) cudaStreamDestroy(h2d_stream_);
        if (d2h_stream_) cudaStreamDestroy(d2h_stream_);
        for (auto& e : ev_) if (e) cudaEventDestroy(e);
        for (auto& pr : ring_) for (auto& e : pr) if (e) cudaEventDestroy(e);
        if (ev_hist_) cudaEventDestroy(ev_hist_);
        if (h_hist_) cudaFreeHost(h_hist_);
        for (auto& sl : slot_)
            for (cudaEvent_t e : {sl.copied, sl.ingested, sl.emitted, sl.drained}) if (e) cudaEventDestroy(e);
        for (auto& t : tickets_) if (t.done) cudaEventDestroy(t.done);
    }

    bool init(const Puncturer* punct, const DecoderOptions& opt) {
        if (opt.device >= 0) LDPC_CUDA_CHECK(cudaSetDevice(opt.device));
        LDPC_CUDA_CHECK(cudaGetDevice(&device_));
        cudaDeviceProp prop;
        LDPC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device_));
        sm_count_ = prop.multiProcessorCount;
        if (punct) punct_ = std::make_unique<Puncturer>(*punct);
        expected_len_ = (size_t)g_.n;
        if (punct_) {
            // reference: depuncture() expands blocks of llrs_len/num_trues; the result must be n
            if (g_.n % (int)punct_->pattern.size() != 0) { set_last_error("codeword size not divisible by puncturing pattern length"); return false; }
            expected_len_ = (size_t)g_.n / punct_->pattern.size() * punct_->num_trues;
            std::vector<int> map;
            if (!punct_->depuncture_map(expected_len_, (size_t)g_.n, &map)) { set_last_error("bad puncturing pattern"); return false; }
            if (!d_src_map_.upload(map)) return false;
        }
        // rows the min* rules cannot process: the reference panics on them during the first
        // iteration (arithmetic.rs:744-745, :952, :971); here such frames report -2.
        panics_ = false;
        if (impl_.rule == Rule::Minstarapprox || impl_.rule == Rule::Aminstar) {
            for (int r = 0; r < g_.m; ++r) {
                int d = g_.row_ptr[(size_t)r + 1] - g_.row_ptr[(size_t)r];
                if (d == 1 || (d == 0 && impl_.rule == Rule::Aminstar)) panics_ = true;
            }
        }
        kind_ = impl_.schedule == Schedule::HorizontalLayered ? Kind::Layered
                : (impl_.dtype == Dtype::I8 ? Kind::FloodI8 : Kind::FloodFloat);
        const int max_deg = kind_ == Kind::FloodI8 ? flood_i8_max_row_degree() : generic_max_row_degree();
        if (g_.max_row_deg > max_deg) {
            set_last_error("row degree above the supported maximum of " + std::to_string(max_deg));
            return false;
        }
        // the packed variable node of K1 sums biased bytes in 16-bit halves: 255*(d+1) must stay below 2^15
        if (kind_ == Kind::FloodI8 && g_.max_col_deg > 126) {
            set_last_error("column degree above the supported maximum of 126 for the int8 flooding decoders");
            return false;
        }
        if (!d_row_ptr_.upload(g_.row_ptr) || !d_col_idx_.upload(g_.col_idx) || !d_col_ptr_.upload(g_.col_ptr) ||
            !d_col_edge_.upload(g_.col_edge))
            return false;
        dg_.n = g_.n; dg_.m = g_.m; dg_.E = g_.E;
        dg_.row_ptr = d_row_ptr_.p; dg_.col_idx = d_col_idx_.p; dg_.col_ptr = d_col_ptr_.p; dg_.col_edge = d_col_edge_.p;
        if (kind_ == Kind::FloodI8 && (!build_row_meta() || !build_var_classes())) return false;
        if (kind_ == Kind::Layered && !build_levels()) return false;
        if (kind_ == Kind::Layered) {
            // K3q (frame per CTA, posteriors in shared memory) when they fit and the level schedule is wide
            int path = opt.layered_path;
            if (const char* e = getenv("LDPC_B200_LAYERED")) path = !strcmp(e, "tile") ? 1 : (!strcmp(e, "smem") ? 2 : path);
            const size_t need = layered_smem_bytes(g_.n, impl_.dtype == Dtype::F64, impl_.dtype == Dtype::I8) + 1024;
            const bool fits = need <= (size_t)prop.sharedMemPerBlockOptin;
            const bool wide = num_levels_ > 0 && g_.m / num_levels_ >= 16;
            use_smem_layered_ = fits && (path == 2 || (path == 0 && wide));
            if (use_smem_layered_ && !build_ell()) return false;
        }
        LDPC_CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
        for (auto& e : ev_) LDPC_CUDA_CHECK(cudaEventCreate(&e));
        for (auto& pr : ring_) { LDPC_CUDA_CHECK(cudaEventCreate(&pr[0])); LDPC_CUDA_CHECK(cudaEventCreate(&pr[1])); }
        ring_ok_ = true;
        max_tiles_opt_ = opt.max_tiles;
        nw_opt_ = opt.words_per_lane;
        if (const char* e = getenv("LDPC_B200_NW")) nw_opt_ = atoi(e);
        if (const char* e = getenv("LDPC_B200_TWO_STAGE")) two_stage_ = atoi(e) != 0;
        if (const char* e = getenv("LDPC_B200_EXACT_LIBM")) exact_libm_ = atoi(e) != 0;
        return true;
    }

    int n() const override { return g_.n; }
    int k() const override { return g_.n - g_.m; }
    int edges() const override { return g_.E; }
    size_t expected_llrs_len() const override { return expected_len_; }
    const BatchStats& stats() const override { return stats_; }

    bool decode(const double* llrs, size_t llrs_len, size_t max_iterations, DecoderOutput* out) override {
        out->codeword.assign((size_t)g_.n, 0);
        int32_t it = 0;
        uint32_t mi = (uint32_t)std::min<size_t>(max_iterations, 0x7fffffffu);
        if (!decode_batch(llrs, true, llrs_len, 1, mi, out->codeword.data(), (size_t)g_.n, (size_t)g_.n, &it)) return false;
        if (it == -2) { set_last_error("check node of degree < 2 with a min* rule (the reference panics)"); return false; }
        out->success = it >= 0;
        out->iterations = it >= 0 ? (size_t)it : max_iterations;
        return true;
    }

    bool decode_batch(const void* llrs, bool is_f64, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                      uint8_t* out, size_t out_len, size_t out_stride, int32_t* iterations) override {
        const int64_t t = submit_batch(llrs, is_f64, llrs_len, nframes, max_iterations, out, out_len, out_stride, iterations);
        return t >= 0 && wait(t);
    }

    // Asynchronous host-buffer decode.  The batch is cut into chunks that flow through a three-stage
    // pipeline — H2D copy (copy engine) -> ingest + BP + emit -> D2H copy — over two staging slots.  The int8
    // flooding decoder on big batches uses chunks of HALF a GPU-filling launch on two compute streams with
    // their own workspaces (a half-launch still fills the GPU: one 2-CTA cluster per tile): the exposed
    // prologue is the copy of half a launch, the other lane's kernels are already queued when a launch ends, and
    // consecutive submit_batch calls keep the pipeline full.  Nothing here blocks the host when the caller's
    // buffers are pinned; wait(ticket) does.
    int64_t submit_batch(const void* llrs, bool is_f64, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                         uint8_t* out, size_t out_len, size_t out_stride, int32_t* iterations) override {
        if (!check_args(llrs_len, out_len, out_stride)) return -1;
        DeviceScope scope(device_);
        if (!scope.ok) { set_last_error("cudaSetDevice failed"); return -1; }
        if (!ensure_pipeline()) return -1;
        const int64_t ticket = next_ticket_++;
        Ticket tk;
        tk.id = ticket;
        if (cudaEventCreateWithFlags(&tk.done, cudaEventDisableTiming) != cudaSuccess) { set_last_error("cudaEventCreate failed"); return -1; }
        if (nframes > 0) {
            const size_t esz = is_f64 ? 8 : 4;
            int lanes = 1;
            const size_t chunk_frames = plan_host_chunks(nframes, 2 * (llrs_len * esz + out_len + 4), &lanes);
            // staging buffers only ever grow; growing them waits for everything in flight first
            const size_t stage_frames = std::min(nframes, chunk_frames);
            const size_t need_in = stage_frames * llrs_len * esz, need_out = std::max<size_t>(stage_frames * out_len, 1);
            const size_t nchunks = (nframes + chunk_frames - 1) / chunk_frames;
            for (int b = 0; b < (nchunks > 1 ? 2 : 1); ++b) {
                StageSlot& sl = slot_[(slot_counter_ + (size_t)b) % 2];
                if (sl.in.count < need_in || sl.out.count < need_out || sl.iters.count < stage_frames) {
                    cudaDeviceSynchronize();
                    if (!sl.in.ensure(need_in) || !sl.out.ensure(need_out) || !sl.iters.ensure(stage_frames)) { cudaEventDestroy(tk.done); return -1; }
                }
            }
            for (size_t f0 = 0; f0 < nframes; f0 += chunk_frames) {
                StageSlot& sl = slot_[slot_counter_ % 2];
                const int w = lanes == 2 ? (int)(slot_counter_ % 2) : 0;
                ++slot_counter_;
                cudaStream_t cs = w == 0 ? stream_ : stream2_;
                const size_t nf = std::min(chunk_frames, nframes - f0);
                if (!enqueue_chunk(sl, ws_[w], cs, (const uint8_t*)llrs + f0 * llrs_len * esz, is_f64, llrs_len, nf, max_iterations,
                                   out + f0 * out_stride, out_len, out_stride, iterations + f0)) {
                    cudaEventDestroy(tk.done);
                    return -1;
                }
            }
        }
        if (cudaEventRecord(tk.done, d2h_stream_) != cudaSuccess) { set_last_error("cudaEventRecord failed"); cudaEventDestroy(tk.done); return -1; }
        tickets_.push_back(tk);
        return ticket;
    }

    bool wait(int64_t ticket) override {
        DeviceScope scope(device_);
        bool ok = true;
        // tickets complete in order (one D2H stream): waiting for one retires every older one too
        while (!tickets_.empty() && tickets_.front().id <= ticket) {
            Ticket tk = tickets_.front();
            tickets_.pop_front();
            cudaError_t e = cudaEventSynchronize(tk.done);
            cudaEventDestroy(tk.done);
            if (e != cudaSuccess) { set_last_error(std::string("decode failed: ") + cudaGetErrorString(e)); ok = false; }
        }
        if (tickets_.empty()) {
            // surface asynchronous kernel errors of the compute streams as well
            for (cudaStream_t st : {stream_, stream2_}) {
                if (!st) continue;
                cudaError_t e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) { set_last_error(std::string("decode failed: ") + cudaGetErrorString(e)); ok = false; }
            }
        }
        return ok;
    }

    // waits for every host-path batch still in flight (no-op when there is none)
    bool drain_tickets() { return tickets_.empty() || wait(tickets_.back().id); }

    // Test hook: decode_batch on one chunk that also returns every frame's posterior LLRs as f64 [nframes][n] — the
    // flooding decoder's output_llrs (flooding.rs:111-125) / the layered decoder's Qv — for the float decoders on the
    // K2 and K3q kernels.  Frames that pass the pre-check (0 iterations) have no posteriors (the reference's are stale).
    bool decode_batch_posteriors(const void* llrs, bool is_f64, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                                 uint8_t* out, size_t out_len, size_t out_stride, int32_t* iterations, double* posteriors) override {
        if (!check_args(llrs_len, out_len, out_stride)) return false;
        if (kind_ == Kind::FloodI8 || impl_.dtype == Dtype::I8 || (kind_ == Kind::Layered && !use_smem_layered_)) {
            set_last_error("posterior output is implemented for the float decoders on the flooding and frame-per-CTA layered kernels");
            return false;
        }
        if (nframes == 0) return true;
        DeviceScope scope(device_);
        if (!scope.ok) { set_last_error("cudaSetDevice failed"); return false; }
        if (nframes > plan_chunk_frames(nframes, 0)) { set_last_error("too many frames for one chunk"); return false; }
        if (!drain_tickets()) return false;                 // this call uses the first lane's stream and workspace
        struct HookReset {                                  // whatever happens below, the next decode must not dump
            GpuDecoder* d;
            ~HookReset() { d->dump_post_tiles_ = nullptr; d->dump_post_ = nullptr; }
        } hook_reset{this};
        const size_t esz = is_f64 ? 8 : 4, n = (size_t)g_.n;
        DevBuf<uint8_t> d_in, d_out, d_post_tiles;
        DevBuf<int32_t> d_it;
        DevBuf<double> d_post;
        const size_t tiles = (nframes + kTileFrames - 1) / kTileFrames;
        if (!d_in.ensure(nframes * llrs_len * esz) || !d_out.ensure(std::max<size_t>(nframes * out_len, 1)) || !d_it.ensure(nframes) ||
            !d_post.ensure(nframes * n))
            return false;
        LDPC_CUDA_CHECK(cudaMemsetAsync(d_post.p, 0, nframes * n * sizeof(double), stream_));
        if (kind_ == Kind::FloodFloat) {
            if (!d_post_tiles.ensure(tiles * n * kTileFrames * elem_size())) return false;
            LDPC_CUDA_CHECK(cudaMemsetAsync(d_post_tiles.p, 0, tiles * n * kTileFrames * elem_size(), stream_));
            dump_post_tiles_ = d_post_tiles.p;
        } else {
            dump_post_ = d_post.p;
        }
        LDPC_CUDA_CHECK(cudaMemcpyAsync(d_in.p, llrs, nframes * llrs_len * esz, cudaMemcpyHostToDevice, stream_));
        const bool ok = run_chunk(ws_[0], d_in.p, is_f64, llrs_len, nframes, max_iterations, d_out.p, out_len, out_len, d_it.p, stream_);
        void* tiles_ptr = dump_post_tiles_;
        if (!ok) return false;
        if (tiles_ptr && !launch_emit_posteriors(tiles_ptr, impl_.dtype == Dtype::F64, g_.n, nframes, d_post.p, stream_)) return false;
        if (out_len) LDPC_CUDA_CHECK(cudaMemcpy2DAsync(out, out_stride, d_out.p, out_len, out_len, nframes, cudaMemcpyDeviceToHost, stream_));
        LDPC_CUDA_CHECK(cudaMemcpyAsync(iterations, d_it.p, nframes * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
        LDPC_CUDA_CHECK(cudaMemcpyAsync(posteriors, d_post.p, nframes * n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        LDPC_CUDA_CHECK(cudaStreamSynchronize(stream_));
        return true;
    }

    bool decode_batch_device(const void* d_llrs, bool is_f64, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                             uint8_t* d_out, size_t out_len, size_t out_stride, int32_t* d_iterations,
                             cudaStream_t stream) override {
        return decode_batch_device_lane(0, d_llrs, is_f64, llrs_len, nframes, max_iterations, d_out, out_len, out_stride, d_iterations, stream);
    }

    bool decode_batch_device_lane(int lane, const void* d_llrs, bool is_f64, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                                  uint8_t* d_out, size_t out_len, size_t out_stride, int32_t* d_iterations,
                                  cudaStream_t stream) override {
        if (lane < 0 || lane > 1) { set_last_error("lane must be 0 or 1"); return false; }
        if (!check_args(llrs_len, out_len, out_stride)) return false;
        if (nframes == 0) return true;
        DeviceScope scope(device_);
        if (!scope.ok) { set_last_error("cudaSetDevice failed"); return false; }
        if (!drain_tickets()) return false;                 // host-path batches in flight own the same workspaces
        const size_t esz = is_f64 ? 8 : 4;
        const size_t chunk_frames = plan_chunk_frames(nframes, 0);
        for (size_t f0 = 0; f0 < nframes; f0 += chunk_frames) {
            size_t nf = std::min(chunk_frames, nframes - f0);
            if (!run_chunk(ws_[lane], (const uint8_t*)d_llrs + f0 * llrs_len * esz, is_f64, llrs_len, nf, max_iterations,
                           d_out + f0 * out_stride, out_len, out_stride, d_iterations + f0, stream))
                return false;
        }
        return true;
    }

private:
    bool check_args(size_t llrs_len, size_t out_len, size_t out_stride) {
        // the reference asserts / panics on these (flooding.rs:56, c_api/decoder.rs:51,:61)
        if (llrs_len != expected_len_) { set_last_error("llrs_len does not match the code"); return false; }
        if (out_len > (size_t)g_.n || out_stride < out_len) { set_last_error("output_len larger than the codeword"); return false; }
        return true;
    }

    // words per lane for a batch: 512-frame tiles (512-byte HBM granules) once they can fill the GPU
    int pick_nw(size_t nframes) const {
        if (kind_ != Kind::FloodI8) return 1;
        if (nw_opt_ == 1 || nw_opt_ == 4) return nw_opt_;
        return nframes >= (size_t)sm_count_ * 512 ? 4 : 1;
    }

    // rules.cuh rule id; LDPC_B200_EXACT_LIBM=1 selects the bit-exact ln(1 + e^-t) for the f32 Min*-approx / A-Min* rules
    int rule_id() const {
        const bool exact = exact_libm_ && impl_.dtype == Dtype::F32;
        switch (impl_.rule) {
            case Rule::Phi: return kPhi;
            case Rule::Tanh: return kTanh;
            case Rule::Minstarapprox: return exact ? kMinstarapproxExact : kMinstarapprox;
            default: return exact ? kAminstarExact : kAminstar;
        }
    }

    // HBM bytes of decoder state per 128 frames
    size_t elem_size() const { return impl_.dtype == Dtype::F64 ? 8 : (impl_.dtype == Dtype::F32 ? 4 : 1); }
    size_t bytes_per_128() const {
        const size_t E = (size_t)g_.E, n = (size_t)g_.n;
        if (kind_ == Kind::FloodI8) return E * kLanes * 5 + n * kLanes * 4 + 2 * n * kLanes + (size_t)g_.m * kLanes;
        const size_t s = elem_size();
        if (kind_ == Kind::FloodFloat) return E * kTileFrames * s + E * kLanes + n * kTileFrames * s + 2 * n * kLanes;
        if (use_smem_layered_) return ell_size_ * kTileFrames * s;
        const size_t qs = impl_.dtype == Dtype::I8 ? 2 : s;
        return E * kTileFrames * s + n * kTileFrames * qs + 2 * n * kLanes;
    }

    size_t held_bytes() const {
        size_t b = ws_[0].bytes() + ws_[1].bytes();
        for (const auto& sl : slot_) b += sl.in.count + sl.out.count;
        return b;
    }

    // frames per kernel launch: whole waves of resident CTAs, bounded by free HBM
    size_t plan_chunk_frames(size_t nframes, size_t staging_bytes_per_frame) {
        // cap in units of 128 frames: two 512-frame tiles (int8 flooding) or four 128-frame tiles per SM
        size_t cap = max_tiles_opt_ > 0 ? (size_t)max_tiles_opt_ : (size_t)sm_count_ * (kind_ == Kind::FloodI8 ? 8 : 4);
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            size_t budget = (size_t)((double)(free_b + held_bytes()) * (staging_bytes_per_frame ? 0.85 : 0.45));
            size_t fit = std::max<size_t>(budget / std::max<size_t>(bytes_per_128() + staging_bytes_per_frame * kTileFrames, 1), 1);
            cap = std::min(cap, fit);
        }
        const size_t need = (nframes + kTileFrames - 1) / kTileFrames;
        // a batch that does not fit one launch is cut into EQUAL chunks (a large chunk plus a small tail would push the
        // tail onto the 128-frame-tile kernel, which streams HBM at half the rate)
        if (need > cap) cap = (need + (need + cap - 1) / cap - 1) / ((need + cap - 1) / cap);
        size_t frames = std::max<size_t>(1, std::min(need, cap)) * kTileFrames;
        const size_t tf = (size_t)kTileFrames * pick_nw(frames);
        if (frames > tf) frames = (need > cap ? (frames + tf - 1) / tf : frames / tf) * tf;   // whole tiles per launch (the last chunk may be ragged)
        else frames = tf;
        return frames;
    }

    // Chunking of a host-buffer batch.  *lanes = 2: chunks of half a GPU-filling launch alternate between two
    // compute streams / workspaces (int8 flooding on 512-frame tiles, batches of more than half a launch).
    size_t plan_host_chunks(size_t nframes, size_t staging_bytes_per_frame, int* lanes) {
        const size_t full = plan_chunk_frames(nframes, staging_bytes_per_frame);
        *lanes = 1;
        if (kind_ != Kind::FloodI8 || pick_nw(full) != 4 || getenv("LDPC_B200_ONE_LANE")) return full;
        const size_t tf = (size_t)kTileFrames * 4;
        const size_t half = full / 2 / tf * tf;
        if (half < (size_t)sm_count_ * tf / 2 || nframes <= half || pick_nw(half) != 4) return full;
        *lanes = 2;
        return half;
    }

    bool ensure_pipeline() {
        if (h2d_stream_) return true;
        LDPC_CUDA_CHECK(cudaStreamCreateWithFlags(&h2d_stream_, cudaStreamNonBlocking));
        LDPC_CUDA_CHECK(cudaStreamCreateWithFlags(&d2h_stream_, cudaStreamNonBlocking));
        LDPC_CUDA_CHECK(cudaStreamCreateWithFlags(&stream2_, cudaStreamNonBlocking));
        for (auto& sl : slot_)
            for (cudaEvent_t* e : {&sl.copied, &sl.ingested, &sl.emitted, &sl.drained}) LDPC_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        return true;
    }

    // H2D of one chunk into its staging slot, ingest + BP + emit on compute stream `cs`, D2H of the results
    bool enqueue_chunk(StageSlot& sl, Workspace& ws, cudaStream_t cs, const void* h_llrs, bool is_f64, size_t llrs_len, size_t nf,
                       uint32_t max_iterations, uint8_t* h_out, size_t out_len, size_t out_stride, int32_t* h_iters) {
        const size_t esz = is_f64 ? 8 : 4;
        if (sl.used) LDPC_CUDA_CHECK(cudaStreamWaitEvent(h2d_stream_, sl.ingested, 0));        // staging-in buffer consumed
        LDPC_CUDA_CHECK(cudaMemcpyAsync(sl.in.p, h_llrs, nf * llrs_len * esz, cudaMemcpyHostToDevice, h2d_stream_));
        LDPC_CUDA_CHECK(cudaEventRecord(sl.copied, h2d_stream_));
        LDPC_CUDA_CHECK(cudaStreamWaitEvent(cs, sl.copied, 0));
        if (sl.used) LDPC_CUDA_CHECK(cudaStreamWaitEvent(cs, sl.drained, 0));                   // staging-out buffer drained
        if (!run_chunk(ws, sl.in.p, is_f64, llrs_len, nf, max_iterations, sl.out.p, out_len, out_len, sl.iters.p, cs, sl.ingested)) return false;
        LDPC_CUDA_CHECK(cudaEventRecord(sl.emitted, cs));
        LDPC_CUDA_CHECK(cudaStreamWaitEvent(d2h_stream_, sl.emitted, 0));
        if (out_len)
            LDPC_CUDA_CHECK(cudaMemcpy2DAsync(h_out, out_stride, sl.out.p, out_len, out_len, nf, cudaMemcpyDeviceToHost, d2h_stream_));
        LDPC_CUDA_CHECK(cudaMemcpyAsync(h_iters, sl.iters.p, nf * sizeof(int32_t), cudaMemcpyDeviceToHost, d2h_stream_));
        LDPC_CUDA_CHECK(cudaEventRecord(sl.drained, d2h_stream_));
        sl.used = true;
        return true;
    }

    // Staircase variables (decoder_impl.hpp, RowMeta): degree 2, last slot of row r and second-to-last slot of
    // row r+1, both rows inside one chunk of chunk_rows_ rows and short enough for K1's register path.
    bool build_row_meta() {
        std::vector<RowMeta> meta((size_t)g_.m);
        std::vector<int> fused_row((size_t)g_.n, -1);
        auto deg_of = [&](int v) { return g_.col_ptr[(size_t)v + 1] - g_.col_ptr[(size_t)v]; };
        auto row_deg = [&](int r) { return g_.row_ptr[(size_t)r + 1] - g_.row_ptr[(size_t)r]; };
        const bool enable = !(getenv("LDPC_B200_FUSE") && atoi(getenv("LDPC_B200_FUSE")) == 0);
        chunk_rows_ = kFuseChunkRowsDefault;
        if (const char* e = getenv("LDPC_B200_FUSE_CHUNK")) {
            const int c = atoi(e);
            if (c >= 1 && c <= 4096 && (c & (c - 1)) == 0) chunk_rows_ = c;
        }
        num_fused_ = 0;
        for (int r = 0; r < g_.m; ++r) {
            const int d = row_deg(r);
            meta[(size_t)r] = RowMeta{g_.row_ptr[(size_t)r], d & 0xffff, -1, d > 0 ? g_.col_idx[(size_t)g_.row_ptr[(size_t)r + 1] - 1] : -1};
        }
        // the kernel addresses the channel LLRs of a fused variable as (row + constant): take the offset of the first
        // candidate (k - 1 for a staircase code) and leave variables that do not follow it to the variable pass
        bool have_off = false;
        fuse_var_off_ = 0;
        for (int r = 0; enable && r + 1 < g_.m; ++r) {
            const int d0 = row_deg(r), d1 = row_deg(r + 1);
            if ((r + 1) % chunk_rows_ == 0 || d0 < 2 || d1 < 2 || d0 > kFuseMaxRowDeg || d1 > kFuseMaxRowDeg) continue;
            const int v = g_.col_idx[(size_t)g_.row_ptr[(size_t)r + 1] - 1];
            if (deg_of(v) != 2 || g_.col_idx[(size_t)g_.row_ptr[(size_t)r + 2] - 2] != v) continue;
            if (d0 == 2 && (meta[(size_t)r].d_flags >> 16 & 1) && g_.col_idx[(size_t)g_.row_ptr[(size_t)r]] == v) continue;   // cannot be both slots of row r
            if (!have_off) { fuse_var_off_ = v - (r + 1); have_off = true; }
            if (v != r + 1 + fuse_var_off_) continue;
            meta[(size_t)r].d_flags |= 1 << 17;
            meta[(size_t)r + 1].d_flags |= 1 << 16;
            meta[(size_t)r + 1].fuse_var = v;
            fused_row[(size_t)v] = r;
            ++num_fused_;
        }
        h_fused_row_ = fused_row;
        // where the stop logic finds a variable's hard decisions (flood_i8.cu)
        std::vector<int> snap((size_t)g_.n);
        for (int v = 0; v < g_.n; ++v) {
            if (fused_row[(size_t)v] >= 0) snap[(size_t)v] = -2 - fused_row[(size_t)v];
            else if (deg_of(v) > 0) snap[(size_t)v] = g_.col_edge[(size_t)g_.col_ptr[(size_t)v]];
            else snap[(size_t)v] = -1;
        }
        if (!d_row_meta_.upload(meta) || !d_fused_row_.upload(snap)) return false;
        return true;
    }

    // variables bucketed by degree so the variable pass runs fixed-degree, unrolled code
    bool build_var_classes() {
        std::vector<int> var_list, var_edges;
        vc_ = VarClasses{};
        int k = 0;
        auto deg_of = [&](int v) { return h_fused_row_[(size_t)v] >= 0 ? 0 : g_.col_ptr[(size_t)v + 1] - g_.col_ptr[(size_t)v]; };   // fused: no class
        for (int d = 1; d <= 8; ++d) {
            int before = (int)var_list.size();
            int ebefore = (int)var_edges.size();
            for (int v = 0; v < g_.n; ++v) {
                if (deg_of(v) != d) continue;
                var_list.push_back(v);
                for (int q = g_.col_ptr[(size_t)v]; q < g_.col_ptr[(size_t)v + 1]; ++q) var_edges.push_back(g_.col_edge[(size_t)q]);
            }
            if ((int)var_list.size() == before) continue;
            vc_.deg[k] = d; vc_.off[k] = before; vc_.edge_off[k] = ebefore;
            ++k;
        }
        int before = (int)var_list.size();
        for (int v = 0; v < g_.n; ++v)
            if (deg_of(v) > 8) var_list.push_back(v);
        if ((int)var_list.size() != before) { vc_.deg[k] = 0; vc_.off[k] = before; vc_.edge_off[k] = 0; ++k; }
        vc_.off[k] = (int)var_list.size();
        vc_.num_classes = k;
        if (!d_var_list_.upload(var_list) || !d_var_edges_.upload(var_edges)) return false;
        vc_.var_list = d_var_list_.p;
        vc_.var_edges = d_var_edges_.p;
        return true;
    }

    // Level schedule for the layered kernels: rows that share no column commute, so a row only has
    // to wait for the earlier rows it shares a column with (horizontal_layered.rs:105-110 order).
    bool build_levels() {
        std::vector<int> col_level((size_t)g_.n, 0), row_level((size_t)g_.m, 0);
        int num_levels = 0;
        for (int r = 0; r < g_.m; ++r) {
            int lv = 0;
            for (int q = g_.row_ptr[(size_t)r]; q < g_.row_ptr[(size_t)r + 1]; ++q) lv = std::max(lv, col_level[(size_t)g_.col_idx[(size_t)q]]);
            row_level[(size_t)r] = lv;
            for (int q = g_.row_ptr[(size_t)r]; q < g_.row_ptr[(size_t)r + 1]; ++q) col_level[(size_t)g_.col_idx[(size_t)q]] = lv + 1;
            num_levels = std::max(num_levels, lv + 1);
        }
        std::vector<int> level_ptr((size_t)num_levels + 1, 0), level_rows((size_t)g_.m);
        for (int r = 0; r < g_.m; ++r) level_ptr[(size_t)row_level[(size_t)r] + 1]++;
        for (int l = 0; l < num_levels; ++l) level_ptr[(size_t)l + 1] += level_ptr[(size_t)l];
        std::vector<int> fill(level_ptr.begin(), level_ptr.end() - 1);
        for (int r = 0; r < g_.m; ++r) level_rows[(size_t)fill[(size_t)row_level[(size_t)r]]++] = r;
        num_levels_ = num_levels;
        h_level_ptr_ = level_ptr;
        h_level_rows_ = level_rows;
        return d_level_ptr_.upload(level_ptr) && d_level_rows_.upload(level_rows);
    }

    // ELL layout of the level schedule for layered_smem.cu
    bool build_ell() {
        std::vector<int> row_deg((size_t)g_.m), ell_col;
        std::vector<int> level_ell((size_t)num_levels_), level_deg((size_t)num_levels_);
        size_t off = 0;
        int widest = 1;
        for (int l = 0; l < num_levels_; ++l) {
            const int r0 = h_level_ptr_[(size_t)l], r1 = h_level_ptr_[(size_t)l + 1], nrows = r1 - r0;
            int maxd = 0;
            for (int i = r0; i < r1; ++i) {
                const int c = h_level_rows_[(size_t)i];
                maxd = std::max(maxd, g_.row_ptr[(size_t)c + 1] - g_.row_ptr[(size_t)c]);
            }
            level_ell[(size_t)l] = (int)off;
            level_deg[(size_t)l] = maxd;
            ell_col.resize(off + (size_t)maxd * nrows, 0);
            for (int i = r0; i < r1; ++i) {
                const int c = h_level_rows_[(size_t)i], e0 = g_.row_ptr[(size_t)c], d = g_.row_ptr[(size_t)c + 1] - e0;
                row_deg[(size_t)i] = d;
                if (d != maxd) level_deg[(size_t)l] = -1;
                for (int j = 0; j < d; ++j) ell_col[off + (size_t)j * nrows + (size_t)(i - r0)] = g_.col_idx[(size_t)e0 + j];
            }
            off += (size_t)maxd * nrows;
            widest = std::max(widest, nrows);
        }
        if (off > 0x7fffffffu) { set_last_error("level schedule too large"); return false; }
        ell_size_ = std::max<size_t>(off, 1);
        smem_threads_ = std::min(kSmemLayeredMaxThreads, std::max(64, (widest + 31) / 32 * 32));
        if (const char* e = getenv("LDPC_B200_K3Q_THREADS")) smem_threads_ = std::min(kSmemLayeredMaxThreads, std::max(32, atoi(e) / 32 * 32));
        if (!d_row_deg_.upload(row_deg) || !d_ell_col_.upload(ell_col) || !d_level_ell_.upload(level_ell) || !d_level_deg_.upload(level_deg))
            return false;
        sg_.n = g_.n; sg_.m = g_.m; sg_.num_levels = num_levels_; sg_.level_ptr = d_level_ptr_.p;
        sg_.row_deg = d_row_deg_.p; sg_.ell_col = d_ell_col_.p; sg_.ell_size = ell_size_;
        sg_.level_ell = d_level_ell_.p; sg_.level_deg = d_level_deg_.p;
        return true;
    }

    bool run_chunk_smem_layered(Workspace& ws, const void* d_llrs, bool is_f64, size_t llrs_len, size_t nf, uint32_t max_it, uint8_t* d_out,
                                size_t out_len, size_t out_stride, int32_t* d_iters, cudaStream_t s, cudaEvent_t after_ingest) {
        const size_t esz = layered_smem_rcv_elem(impl_.dtype == Dtype::F64, impl_.dtype == Dtype::I8);
        if (!ws.msg.ensure(nf * ell_size_ * esz)) return false;
        cudaEventRecord(ev_[0], s);
        cudaEventRecord(ev_[1], s);
        LayeredSmemLaunch L{};
        L.graph = sg_;
        L.rule = rule_id();
        L.is_f64 = impl_.dtype == Dtype::F64; L.is_i8 = impl_.dtype == Dtype::I8; L.hardlimit = impl_.hardlimit;
        L.threads = smem_threads_;
        L.llrs = d_llrs; L.in_f64 = is_f64; L.llrs_len = llrs_len; L.nframes = nf; L.src_map = punct_ ? d_src_map_.p : nullptr;
        L.rcv = ws.msg.p; L.out = d_out; L.out_len = out_len; L.out_stride = out_stride; L.iters = d_iters;
        L.max_iter = panics_ ? 0 : (int)std::min<uint32_t>(max_it, 0x7ffffff0u);
        L.post = dump_post_;
        if (!launch_layered_smem(L, s)) return false;
        if (after_ingest) cudaEventRecord(after_ingest, s);     // the caller's LLR buffer is consumed by the decode kernel itself
        cudaEventRecord(ev_[2], s);
        if (panics_ && max_it > 0) {
            if (!launch_mark_panics(d_iters, nf, s)) return false;
        }
        cudaEventRecord(ev_[3], s);
        stats_.kernel_launches += 1;
        timed_ = true;
        return true;
    }

    bool ensure_workspace(Workspace& ws, size_t tiles, int nw) {
        const size_t E = (size_t)g_.E, n = (size_t)g_.n;
        size_t msg_b, hbit_b, in_b, bits_b;
        if (kind_ == Kind::FloodI8) {
            const size_t hb = nw == 4 ? 2 : 1;
            msg_b = tiles * E * kLanes * nw * 4; hbit_b = tiles * E * kLanes * hb; in_b = tiles * n * kLanes * nw * 4; bits_b = tiles * n * kLanes * hb;
        } else if (kind_ == Kind::FloodFloat) {
            msg_b = tiles * E * kTileFrames * elem_size(); hbit_b = tiles * E * kLanes; in_b = tiles * n * kTileFrames * elem_size(); bits_b = tiles * n * kLanes;
        } else {
            const size_t qs = impl_.dtype == Dtype::I8 ? 2 : elem_size();
            msg_b = tiles * E * kTileFrames * elem_size(); hbit_b = 1; in_b = tiles * n * kTileFrames * qs; bits_b = tiles * n * kLanes;
        }
        if (!ws.msg.ensure(std::max<size_t>(msg_b, 1)) || !ws.hbit.ensure(std::max<size_t>(hbit_b, 1)) || !ws.inq.ensure(std::max<size_t>(in_b, 1)) ||
            !ws.hard.ensure(std::max<size_t>(bits_b, 1)) || !ws.final_hard.ensure(std::max<size_t>(bits_b, 1)))
            return false;
        if (kind_ == Kind::FloodI8 && !ws.cbit.ensure(std::max<size_t>(tiles * 2 * (size_t)g_.m * kLanes * (nw == 4 ? 2 : 1), 1))) return false;
        return true;
    }

    // One ingest -> BP kernel -> emit pass over nf frames.  ev_begin / ev_end select which of the four timing
    // events this pass records (a two-stage chunk spreads them over its passes).
    bool run_pass(Workspace& ws, const void* d_llrs, bool is_f64, size_t llrs_len, size_t nf, uint32_t max_it, uint8_t* d_out,
                  size_t out_len, size_t out_stride, int32_t* d_iters, cudaStream_t s, cudaEvent_t after_ingest = nullptr,
                  bool ev_begin = true, bool ev_end = true) {
        if (use_smem_layered_) return run_chunk_smem_layered(ws, d_llrs, is_f64, llrs_len, nf, max_it, d_out, out_len, out_stride, d_iters, s, after_ingest);
        const int nw = pick_nw(nf);
        const size_t tf = (size_t)kTileFrames * nw;
        const int tiles = (int)((nf + tf - 1) / tf);
        if (!ensure_workspace(ws, (size_t)tiles, nw)) return false;
        if (!ws.iters_tile.ensure((size_t)tiles * tf)) return false;
        if (ev_begin) cudaEventRecord(ev_[0], s);
        IngestLaunch in{};
        in.llrs = d_llrs; in.is_f64 = is_f64; in.llrs_len = llrs_len; in.nframes = nf; in.n = g_.n;
        in.src_map = punct_ ? d_src_map_.p : nullptr; in.num_tiles = tiles; in.words_per_lane = nw;
        in.hard = ws.hard.p;
        if (kind_ == Kind::FloodI8) in.inq_i8 = reinterpret_cast<uint32_t*>(ws.inq.p);
        else if (impl_.dtype == Dtype::F32) in.in_f32 = reinterpret_cast<float*>(ws.inq.p);
        else if (impl_.dtype == Dtype::F64) in.in_f64 = reinterpret_cast<double*>(ws.inq.p);
        else in.in_i16 = reinterpret_cast<int16_t*>(ws.inq.p);
        if (!launch_ingest(in, s)) return false;
        if (ev_begin) cudaEventRecord(ev_[1], s);
        if (ev_begin && ev_end && ring_ok_) cudaEventRecord(ring_[ring_count_ % kRing][0], s);
        if (after_ingest) cudaEventRecord(after_ingest, s);
        // a graph the min* rules panic on: run only the pre-check; everything else reports -2
        const int max_iter = panics_ ? 0 : (int)std::min<uint32_t>(max_it, 0x7ffffff0u);
        if (kind_ == Kind::FloodI8) {
            FloodI8Launch fl{};
            fl.graph = dg_; fl.classes = vc_; fl.num_tiles = tiles; fl.words_per_lane = nw;
            fl.msg = reinterpret_cast<uint32_t*>(ws.msg.p); fl.hbit = ws.hbit.p; fl.inq = reinterpret_cast<const uint32_t*>(ws.inq.p);
            fl.raw0 = ws.hard.p; fl.final_hard = ws.final_hard.p; fl.iters = ws.iters_tile.p; fl.max_iter = max_iter;
            fl.row_meta = d_row_meta_.p; fl.snap_src = d_fused_row_.p; fl.snap_n = (int)std::min<size_t>(out_len, (size_t)g_.n); fl.cbit = ws.cbit.p; fl.chunk_rows = chunk_rows_; fl.fuse_var_off = fuse_var_off_; fl.graph_max_row_deg = g_.max_row_deg;
            fl.aminstar = impl_.rule == Rule::Aminstar; fl.jones = impl_.jones; fl.hardlimit = impl_.hardlimit; fl.deg1clip = impl_.deg1clip;
            // rows beyond the register path (8 edges; 10 on 128-frame tiles) are folded from a wide shared-memory stage
            const int reg_cap = nw == 4 ? 8 : 10;
            fl.wide_cap = g_.max_row_deg <= reg_cap ? 0 : (g_.max_row_deg <= 16 ? 16 : 32);
            if (const char* e = getenv("LDPC_B200_WIDE")) fl.wide_cap = atoi(e);
            fl.cluster = 1;
            while (fl.cluster < 16 && tiles * fl.cluster * 2 <= 2 * sm_count_) fl.cluster *= 2;     // up to ~2 CTAs per SM
            if (const char* e = getenv("LDPC_B200_CLUSTER")) fl.cluster = atoi(e);
            if (!launch_flood_i8(fl, s)) return false;
        } else {
            GenericLaunch gl{};
            gl.graph = dg_; gl.num_tiles = tiles; gl.is_f64 = impl_.dtype == Dtype::F64; gl.is_i8 = impl_.dtype == Dtype::I8;
            gl.hardlimit = impl_.hardlimit;
            gl.rule = rule_id();
            gl.msg = ws.msg.p; gl.hbit = ws.hbit.p; gl.in = ws.inq.p; gl.in_out_q = ws.inq.p;
            gl.raw0 = ws.hard.p; gl.final_hard = ws.final_hard.p; gl.iters = ws.iters_tile.p; gl.max_iter = max_iter;
            gl.level_ptr = d_level_ptr_.p; gl.level_rows = d_level_rows_.p; gl.num_levels = num_levels_;
            gl.post = dump_post_tiles_;
            gl.cluster = 1;
            while (gl.cluster < 8 && tiles * gl.cluster * 2 <= 2 * sm_count_) gl.cluster *= 2;      // up to ~2 CTAs per SM
            if (const char* e = getenv("LDPC_B200_CLUSTER")) gl.cluster = atoi(e);
            if (!(kind_ == Kind::FloodFloat ? launch_flood_float(gl, s) : launch_layered(gl, s))) return false;
        }
        if (ev_end) cudaEventRecord(ev_[2], s);
        if (ev_begin && ev_end && ring_ok_) {               // averaged BP-kernel time without a host sync per call (resolve_stats)
            const size_t slot = ring_count_ % kRing;
            if (ring_count_ - ring_resolved_ >= kRing) ++ring_resolved_;        // oldest unresolved entry is overwritten
            cudaEventRecord(ring_[slot][1], s);
            ring_pending_[slot] = true;
            ++ring_count_;
        }
        EmitLaunch em{};
        em.final_hard = ws.final_hard.p; em.n = g_.n; em.num_tiles = tiles; em.words_per_lane = nw; em.nframes = nf; em.out = d_out; em.out_len = out_len;
        em.out_stride = out_stride;
        if (!launch_emit(em, s)) return false;
        LDPC_CUDA_CHECK(cudaMemcpyAsync(d_iters, ws.iters_tile.p, nf * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
        if (panics_ && max_it > 0) {
            if (!launch_mark_panics(d_iters, nf, s)) return false;
        }
        if (ev_end) cudaEventRecord(ev_[3], s);
        stats_.kernel_launches += 3;
        timed_ = true;
        return true;
    }

    // ---- straggler re-decode ----------------------------------------------------------------------
    // A tile runs until its slowest frame stops, so at operating points where most frames converge
    // early a few stragglers make every tile pay max_iterations.  When the iteration histogram of the
    // previous chunk says it pays off, a chunk is decoded in two stages: every frame for at most m1
    // iterations, then the frames still unconverged are gathered into a small dense batch and decoded
    // again FROM THEIR LLRs with the full iteration budget.  Decoding is deterministic and frames are
    // independent, so words and iteration counts are exactly those of a single pass.
    uint32_t two_stage_m1(uint32_t max_it, size_t nf) {
        if (!two_stage_ || panics_ || use_smem_layered_ || kind_ == Kind::Layered || max_it < 8 || max_it > 127) return 0;
        const size_t tf = (size_t)kTileFrames * pick_nw(nf);
        if (nf < 8 * tf || !hist_valid_ || hist_max_it_ != max_it) return 0;
        cudaEventSynchronize(ev_hist_);
        double total = 0;
        for (uint32_t b = 0; b <= max_it; ++b) total += (double)h_hist_[b];
        if (total < 1) return 0;
        // gt[M] = fraction of frames that need more than M iterations (failures sit in bin max_it)
        double best_cost = 1e300, tail = 0, single = (double)max_it;
        uint32_t best_m = 0;
        std::vector<double> gt((size_t)max_it + 1, 0.0);
        for (uint32_t m = max_it; m-- > 0;) { tail += (double)h_hist_[m + 1]; gt[m] = tail / total; }
        for (uint32_t m = 1; m < max_it; ++m) {
            if (gt[m] < 0.5 / (double)tf && single == (double)max_it) single = (double)m;     // a tile's slowest frame, roughly
            const double cost = (double)m + gt[m] * (double)max_it;
            if (cost < best_cost) { best_cost = cost; best_m = m; }
        }
        return best_cost < 0.9 * single ? best_m : 0;
    }

    bool run_chunk(Workspace& ws, const void* d_llrs, bool is_f64, size_t llrs_len, size_t nf, uint32_t max_it, uint8_t* d_out,
                   size_t out_len, size_t out_stride, int32_t* d_iters, cudaStream_t s, cudaEvent_t after_ingest = nullptr) {
        const uint32_t m1 = two_stage_m1(max_it, nf);
        if (m1 == 0) {
            if (!run_pass(ws, d_llrs, is_f64, llrs_len, nf, max_it, d_out, out_len, out_stride, d_iters, s, after_ingest)) return false;
            return record_histogram(d_iters, nf, max_it, s);
        }
        const size_t esz = is_f64 ? 8 : 4;
        if (!run_pass(ws, d_llrs, is_f64, llrs_len, nf, m1, d_out, out_len, out_stride, d_iters, s, nullptr, true, false)) return false;
        if (!d_fail_idx_.ensure(nf + 1)) return false;
        LDPC_CUDA_CHECK(cudaMemsetAsync(d_fail_idx_.p + nf, 0, sizeof(int32_t), s));
        if (!launch_collect_failed(d_iters, nf, d_fail_idx_.p, d_fail_idx_.p + nf, s)) return false;
        int32_t count = 0;
        LDPC_CUDA_CHECK(cudaMemcpyAsync(&count, d_fail_idx_.p + nf, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        LDPC_CUDA_CHECK(cudaStreamSynchronize(s));
        const size_t tf = (size_t)kTileFrames * pick_nw(nf);
        const size_t cap = std::max(tf, (nf / 4 + tf - 1) / tf * tf);               // stage-2 sub-batch, whole tiles
        if (count > 0) {
            const size_t sub = std::min<size_t>(cap, (size_t)count);
            if (!d_llrs2_.ensure(sub * llrs_len * esz) || !d_out2_.ensure(std::max<size_t>(sub * out_len, 1)) || !d_iters2_.ensure(sub)) return false;
        }
        for (size_t off = 0; off < (size_t)count; off += cap) {
            const size_t c = std::min(cap, (size_t)count - off);
            const bool last = off + c >= (size_t)count;
            if (!launch_gather_rows(d_llrs, llrs_len * esz, d_fail_idx_.p + off, c, d_llrs2_.p, s)) return false;
            if (!run_pass(ws, d_llrs2_.p, is_f64, llrs_len, c, max_it, d_out2_.p, out_len, out_len, d_iters2_.p, s, nullptr, false, last)) return false;
            if (!launch_scatter_results(d_out2_.p, out_len, d_iters2_.p, d_fail_idx_.p + off, c, d_out, out_stride, d_iters, s)) return false;
        }
        if (count == 0) { cudaEventRecord(ev_[2], s); cudaEventRecord(ev_[3], s); }
        if (after_ingest) cudaEventRecord(after_ingest, s);        // the caller's LLRs were still needed by the gathers
        stats_.kernel_launches += 1 + (count > 0 ? 2 : 0);
        ++two_stage_chunks_;
        return record_histogram(d_iters, nf, max_it, s);
    }

    bool record_histogram(const int32_t* d_iters, size_t nf, uint32_t max_it, cudaStream_t s) {
        if (!two_stage_ || use_smem_layered_ || kind_ == Kind::Layered || max_it > 127) return true;
        if (!d_hist_.ensure(128)) return false;
        if (!h_hist_) {
            LDPC_CUDA_CHECK(cudaMallocHost(&h_hist_, 128 * sizeof(unsigned int)));
            LDPC_CUDA_CHECK(cudaEventCreateWithFlags(&ev_hist_, cudaEventDisableTiming));
        }
        LDPC_CUDA_CHECK(cudaMemsetAsync(d_hist_.p, 0, 128 * sizeof(unsigned int), s));
        if (!launch_iter_histogram(d_iters, nf, max_it, d_hist_.p, s)) return false;
        LDPC_CUDA_CHECK(cudaMemcpyAsync(h_hist_, d_hist_.p, 128 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
        LDPC_CUDA_CHECK(cudaEventRecord(ev_hist_, s));
        hist_valid_ = true;
        hist_max_it_ = max_it;
        return true;
    }

    bool launch_mark_panics(int32_t* d_iters, size_t nf, cudaStream_t s);
    bool launch_collect_failed(const int32_t* d_iters, size_t nf, int32_t* idx, int32_t* count, cudaStream_t s);
    bool launch_gather_rows(const void* src, size_t row_bytes, const int32_t* idx, size_t count, void* dst, cudaStream_t s);
    bool launch_scatter_results(const uint8_t* out2, size_t out_len, const int32_t* iters2, const int32_t* idx, size_t count,
                                uint8_t* out, size_t out_stride, int32_t* iters, cudaStream_t s);
    bool launch_iter_histogram(const int32_t* d_iters, size_t nf, uint32_t max_it, unsigned int* hist, cudaStream_t s);

public:
    // resolves the event timings of the last chunk (synchronises the stream)
    // average BP-kernel time of the passes since the previous call (up to kRing of them), and how many
    float average_decode_ms(int64_t* launches) {
        DeviceScope scope(device_);
        double sum = 0;
        int64_t n = 0;
        for (; ring_resolved_ < ring_count_; ++ring_resolved_) {
            const size_t slot = ring_resolved_ % kRing;
            if (!ring_pending_[slot]) continue;
            float ms = 0;
            if (cudaEventSynchronize(ring_[slot][1]) == cudaSuccess && cudaEventElapsedTime(&ms, ring_[slot][0], ring_[slot][1]) == cudaSuccess) { sum += ms; ++n; }
            ring_pending_[slot] = false;
        }
        if (launches) *launches = n;
        return n ? (float)(sum / (double)n) : 0.0f;
    }

    void resolve_stats() {
        if (!timed_) return;
        cudaEventSynchronize(ev_[3]);
        cudaEventElapsedTime(&stats_.ingest_ms, ev_[0], ev_[1]);
        cudaEventElapsedTime(&stats_.decode_ms, ev_[1], ev_[2]);
        cudaEventElapsedTime(&stats_.emit_ms, ev_[2], ev_[3]);
    }

private:
    DecoderImplementation impl_;
    Graph g_;
    DeviceGraph dg_{};
    std::unique_ptr<Puncturer> punct_;
    size_t expected_len_ = 0;
    bool panics_ = false;
    int device_ = 0, sm_count_ = 148, max_tiles_opt_ = 0, nw_opt_ = 0;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
    static constexpr size_t kRing = 32;
    cudaEvent_t ring_[kRing][2] = {};
    bool ring_pending_[kRing] = {};
    bool ring_ok_ = false;
    size_t ring_count_ = 0, ring_resolved_ = 0;
    bool timed_ = false;
    BatchStats stats_;
    enum class Kind { FloodI8, FloodFloat, Layered };
    Kind kind_ = Kind::FloodI8;
    DevBuf<int> d_row_ptr_, d_col_idx_, d_col_ptr_, d_col_edge_, d_src_map_, d_var_list_, d_var_edges_, d_level_ptr_, d_level_rows_;
    DevBuf<RowMeta> d_row_meta_;
    DevBuf<int> d_fused_row_;
    std::vector<int> h_fused_row_;
    int num_fused_ = 0, chunk_rows_ = kFuseChunkRowsDefault, fuse_var_off_ = 0;
    VarClasses vc_{};
    int num_levels_ = 0;
    bool use_smem_layered_ = false;
    std::vector<int> h_level_ptr_, h_level_rows_;
    DevBuf<int> d_row_deg_, d_ell_col_, d_level_ell_, d_level_deg_;
    LayeredSmemGraph sg_{};
    size_t ell_size_ = 1;
    int smem_threads_ = 256;
    void* dump_post_tiles_ = nullptr;  // test hook (decode_batch_posteriors): K2 posterior tiles
    double* dump_post_ = nullptr;      // test hook: K3q posteriors [nframes][n]
    Workspace ws_[2];                  // [0]: device-buffer path and lane 0 of the host path; [1]: lane 1
    StageSlot slot_[2];
    size_t slot_counter_ = 0;
    struct Ticket { int64_t id = 0; cudaEvent_t done = nullptr; };
    std::deque<Ticket> tickets_;
    int64_t next_ticket_ = 0;
    cudaStream_t stream2_ = nullptr;
    DevBuf<int32_t> d_fail_idx_, d_iters2_;
    DevBuf<uint8_t> d_llrs2_, d_out2_;
    DevBuf<unsigned int> d_hist_;
    unsigned int* h_hist_ = nullptr;
    cudaEvent_t ev_hist_ = nullptr;
    bool exact_libm_ = false;
    bool two_stage_ = false, hist_valid_ = false;      // opt-in (LDPC_B200_TWO_STAGE=1): exact, but measured no faster on DVB-S2
    uint32_t hist_max_it_ = 0;
    long long two_stage_chunks_ = 0;
    cudaStream_t h2d_stream_ = nullptr, d2h_stream_ = nullptr;
};

__global__ void mark_panics_kernel(int32_t* iters, size_t nf) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nf && iters[i] < 0) iters[i] = -2;
}

// frames that did not converge within the first stage, in arbitrary order
__global__ void collect_failed_kernel(const int32_t* iters, size_t nf, int32_t* idx, int32_t* count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool f = i < nf && iters[i] < 0;
    const unsigned m = __ballot_sync(0xffffffffu, f);
    int base = 0;
    const int lane = threadIdx.x & 31;
    if (lane == 0 && m) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (f) idx[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
}

__global__ void gather_rows_kernel(const uint32_t* src, size_t row_words, const int32_t* idx, uint32_t* dst) {
    const uint32_t* s = src + (size_t)idx[blockIdx.x] * row_words;
    uint32_t* d = dst + (size_t)blockIdx.x * row_words;
    for (size_t w = threadIdx.x; w < row_words; w += blockDim.x) d[w] = s[w];
}

__global__ void scatter_results_kernel(const uint8_t* out2, size_t out_len, const int32_t* iters2, const int32_t* idx, uint8_t* out,
                                       size_t out_stride, int32_t* iters) {
    const size_t f = (size_t)idx[blockIdx.x];
    const uint8_t* s = out2 + (size_t)blockIdx.x * out_len;
    uint8_t* d = out + f * out_stride;
    for (size_t b = threadIdx.x; b < out_len; b += blockDim.x) d[b] = s[b];
    if (threadIdx.x == 0) iters[f] = iters2[blockIdx.x];
}

__global__ void iter_histogram_kernel(const int32_t* iters, size_t nf, uint32_t max_it, unsigned int* hist) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    const int32_t it = iters[i];
    atomicAdd(&hist[it < 0 || (uint32_t)it > max_it ? max_it : (uint32_t)it], 1u);
}

bool GpuDecoder::launch_collect_failed(const int32_t* d_iters, size_t nf, int32_t* idx, int32_t* count, cudaStream_t s) {
    collect_failed_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, s>>>(d_iters, nf, idx, count);
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

bool GpuDecoder::launch_gather_rows(const void* src, size_t row_bytes, const int32_t* idx, size_t count, void* dst, cudaStream_t s) {
    gather_rows_kernel<<<(unsigned)count, 256, 0, s>>>(static_cast<const uint32_t*>(src), row_bytes / 4, idx, static_cast<uint32_t*>(dst));
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

bool GpuDecoder::launch_scatter_results(const uint8_t* out2, size_t out_len, const int32_t* iters2, const int32_t* idx, size_t count,
                                        uint8_t* out, size_t out_stride, int32_t* iters, cudaStream_t s) {
    scatter_results_kernel<<<(unsigned)count, 256, 0, s>>>(out2, out_len, iters2, idx, out, out_stride, iters);
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

bool GpuDecoder::launch_iter_histogram(const int32_t* d_iters, size_t nf, uint32_t max_it, unsigned int* hist, cudaStream_t s) {
    iter_histogram_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, s>>>(d_iters, nf, max_it, hist);
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

bool GpuDecoder::launch_mark_panics(int32_t* d_iters, size_t nf, cudaStream_t s) {
    mark_panics_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, s>>>(d_iters, nf);
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

}  // namespace

std::unique_ptr<LdpcDecoder> build_decoder(const DecoderImplementation& impl, const Graph& h, const Puncturer* puncturer,
                                           const DecoderOptions& opt) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_last_error("no CUDA device: this library has no CPU fallback");
        return nullptr;
    }
    auto d = std::make_unique<GpuDecoder>(impl, h);
    if (!d->init(puncturer, opt)) return nullptr;
    return d;
}

void resolve_decoder_stats(LdpcDecoder* d) { static_cast<GpuDecoder*>(d)->resolve_stats(); }
float average_decode_ms(LdpcDecoder* d, int64_t* launches) { return static_cast<GpuDecoder*>(d)->average_decode_ms(launches); }

}  // namespace ldpc
