// ldpc_toolbox_b200/csrc/ber.cu — K4/K5: the BER Monte-Carlo loop on the device.
//
// Replaces Worker::simulate (reference src/simulation/ber.rs:436-481) for BPSK/AWGN:
//   random_message            ber.rs:483-488     -> Philox4x32-10 bits, counter = (global frame, word)
//   Encoder::encode           src/encoder.rs:99-120 -> bit-packed staircase scan / packed dense G0
//   Puncturer::puncture       src/simulation/puncturing.rs:47-75
//   BpskModulator::modulate   src/simulation/modulation.rs:87-95   (bit 0 -> -1, bit 1 -> +1)
//   AwgnChannel::add_noise    src/simulation/channel.rs:60-72      -> Philox + Box-Muller
//   BpskDemodulator           src/simulation/modulation.rs:123-141 (LLR = -2 y / sigma^2)
//   decode + error counting   ber.rs:462-474, statistics ber.rs:313-337
// Only the counters return to the host.  Frames are identified by a global index, so results do
// not depend on the batch size or on how frames are sharded over GPUs.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "ber.hpp"
#include "decoder_impl.hpp"
#include "device_common.cuh"

namespace ldpc {
namespace {

// ---- Philox4x32-10 (Salmon et al., SC'11) ------------------------------------------------------
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

constexpr uint32_t kStreamMessage = 0x6d736701u, kStreamNoise = 0x6e6f6902u;

struct FrontendParams {
    int n, m, k, n_tx;
    int staircase;
    const int* h0_ptr;          // staircase: CSR of H0
    const int* h0_idx;
    const uint32_t* g0;         // dense: m x words32 packed rows of G0 = H1^-1 H0
    int g0_words;
    const int* kept;            // transmitted codeword positions (null: all)
    uint64_t first_frame;       // global index of frame 0 of this launch
    uint32_t seed_lo, seed_hi;
    float sigma, llr_scale;     // llr = llr_scale * y, llr_scale = -2/sigma^2
    int modulation;             // 0: BPSK, 1: 8PSK (DVB-S2 Gray mapping)
    int il_cols, il_backwards;  // bit interleaver: columns (0 = none), rows read backwards
    double sigma_d;             // 8PSK: noise sigma in f64 (demapper scale 1/sigma^2)
    float* llrs;                // [nframes][n_tx]
    uint32_t* messages;         // [nframes][ceil(k/32)]
};

// XOR prefix over `nbits` bits held as words in shared memory (in place, inclusive)
__device__ void prefix_xor_bits(uint32_t* w, int nwords) {
    __shared__ uint32_t s_chunk[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nwords; base += 1024) {           // 32 chunks of 32 words
        for (int ch = warp; ch < 32; ch += nwarps) {
            int i = base + ch * 32 + lane;
            uint32_t x = i < nwords ? w[i] : 0;
            x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;
            uint32_t carry = x >> 31;                            // parity of the word
            uint32_t inc = carry;                                // inclusive XOR scan of carries in the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc ^= t;
            }
            uint32_t exc = inc ^ carry;
            if (exc) x = ~x;
            if (i < nwords) w[i] = x;
            if (lane == 31) s_chunk[ch] = inc;
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t c = s_chunk[lane], inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc ^= t;
            }
            uint32_t carry_in = s_carry;
            s_chunk[lane] = (inc ^ c) ^ carry_in;                // exclusive prefix incl. previous super-chunks
            __syncwarp();
            if (lane == 31) s_carry = inc ^ carry_in;
        }
        __syncthreads();
        for (int ch = warp; ch < 32; ch += nwarps) {
            int i = base + ch * 32 + lane;
            if (i < nwords && s_chunk[ch]) w[i] = ~w[i];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) ber_frontend_kernel(FrontendParams p) {
    extern __shared__ uint32_t smem[];
    const int kw = (p.k + 31) / 32, mw = (p.m + 31) / 32;
    uint32_t* msg = smem;              // k bits
    uint32_t* par = smem + kw;         // m bits
    const uint64_t frame = p.first_frame + blockIdx.x;
    const Philox rng{p.seed_lo, p.seed_hi};
    const uint32_t f_lo = (uint32_t)frame, f_hi = (uint32_t)(frame >> 32);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    // ---- random message (uniform bits)
    for (int w4 = threadIdx.x; w4 * 4 < kw; w4 += blockDim.x) {
        uint4 r = rng(f_lo, f_hi, (uint32_t)w4, kStreamMessage);
        uint32_t v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int w = w4 * 4 + i;
            if (w < kw) {
                uint32_t x = v[i];
                if (w == kw - 1 && (p.k & 31)) x &= (1u << (p.k & 31)) - 1u;
                msg[w] = x;
                p.messages[(size_t)blockIdx.x * kw + w] = x;
            }
        }
    }
    __syncthreads();

    // ---- systematic encode: parity bits
    if (p.staircase) {
        for (int j0 = warp * 32; j0 < p.m; j0 += nwarps * 32) {
            int j = j0 + lane;
            uint32_t s = 0;
            if (j < p.m)
                for (int q = __ldg(p.h0_ptr + j); q < __ldg(p.h0_ptr + j + 1); ++q) {
                    int c = __ldg(p.h0_idx + q);
                    s ^= msg[c >> 5] >> (c & 31);
                }
            uint32_t word = __ballot_sync(0xffffffffu, s & 1u);
            if (lane == 0) par[j0 >> 5] = word;
        }
        __syncthreads();
        prefix_xor_bits(par, mw);                                  // accumulate, encoder.rs:112-116
    } else {
        for (int j0 = warp * 32; j0 < p.m; j0 += nwarps * 32) {
            uint32_t word = 0;
            for (int b = 0; b < 32 && j0 + b < p.m; ++b) {
                const uint32_t* row = p.g0 + (size_t)(j0 + b) * p.g0_words;
                uint32_t acc = 0;
                for (int w = lane; w < kw; w += 32) acc ^= __ldg(row + w) & msg[w];
                acc = __popc(acc) & 1u;
#pragma unroll
                for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
                word |= acc << b;
            }
            if (lane == 0) par[j0 >> 5] = word;
        }
        __syncthreads();
    }

    float* out = p.llrs + (size_t)blockIdx.x * p.n_tx;
    auto tx_bit = [&](int t) -> uint32_t {                       // t-th transmitted (punctured) code bit
        const int v = p.kept ? __ldg(p.kept + t) : t;
        return v < p.k ? (msg[v >> 5] >> (v & 31)) & 1u : (par[(v - p.k) >> 5] >> ((v - p.k) & 31)) & 1u;
    };
    if (p.modulation == 1) {
        // ---- 8PSK: interleave (interleaving.rs:40-58), map (modulation.rs:166-178), complex AWGN
        // (channel.rs:74-81), exact max* demapper (modulation.rs:222-262), deinterleave (:64-85).
        // One symbol per thread; the LLR of interleaved position i goes back to transmitted position t(i).
        const int nsym = p.n_tx / 3, rows = p.il_cols > 0 ? p.n_tx / p.il_cols : 0;
        auto src_of = [&](int i) {                               // transmitted index read at interleaved position i
            if (p.il_cols <= 0) return i;
            const int r = i / p.il_cols, c = i % p.il_cols;
            return (p.il_backwards ? p.il_cols - 1 - c : c) * rows + r;
        };
        const double a = 0.70710678118654757, scale = 1.0 / (p.sigma_d * p.sigma_d);
        auto maxstar = [](double x, double y) { return fmax(x, y) + log1p(exp(-fabs(x - y))); };
        for (int sidx = threadIdx.x; sidx < nsym; sidx += blockDim.x) {
            const int t0 = src_of(3 * sidx), t1 = src_of(3 * sidx + 1), t2 = src_of(3 * sidx + 2);
            const uint32_t b0 = tx_bit(t0), b1 = tx_bit(t1), b2 = tx_bit(t2);
            // (b0, b1, b2) -> point; index = b0 | b1 << 1 | b2 << 2
            const double re_tab[8] = {a, 0.0, -1.0, -a, 1.0, a, -a, 0.0};      // 000 100 010 110 001 101 011 111
            const double im_tab[8] = {a, 1.0, 0.0, a, 0.0, -a, -a, -1.0};
            const int idx = (int)(b0 | b1 << 1 | b2 << 2);
            uint4 r = rng(f_lo, f_hi, (uint32_t)sidx, kStreamNoise);
            const double rad = sqrt(-2.0 * log(((double)r.x + 0.5) * 2.3283064365386963e-10));
            double sn, cs;
            sincospi(2.0 * ((double)r.y + 0.5) * 2.3283064365386963e-10, &sn, &cs);
            const double yr = (re_tab[idx] + p.sigma_d * rad * cs) * scale, yi = (im_tab[idx] + p.sigma_d * rad * sn) * scale;
            const double d000 = yr * a + yi * a, d100 = yi, d110 = -yr * a + yi * a, d010 = -yr, d011 = -yr * a - yi * a, d111 = -yi,
                         d101 = yr * a - yi * a, d001 = yr;
            const double l0 = maxstar(maxstar(maxstar(d000, d001), d010), d011) - maxstar(maxstar(maxstar(d100, d101), d110), d111);
            const double l1 = maxstar(maxstar(maxstar(d000, d001), d100), d101) - maxstar(maxstar(maxstar(d010, d011), d110), d111);
            const double l2 = maxstar(maxstar(maxstar(d000, d010), d100), d110) - maxstar(maxstar(maxstar(d001, d011), d101), d111);
            out[t0] = (float)l0; out[t1] = (float)l1; out[t2] = (float)l2;
        }
        return;
    }
    // ---- puncture, BPSK, AWGN, demodulate: four transmitted symbols per thread and Philox call
    // (a bit interleaver in front of BPSK only renames i.i.d. noise samples, so it is the identity here)
    for (int t4 = threadIdx.x; t4 * 4 < p.n_tx; t4 += blockDim.x) {
        uint4 r = rng(f_lo, f_hi, (uint32_t)t4, kStreamNoise);
        // Box-Muller on (u1, u2) pairs.  u1 = (x + 0.5) 2^-32 keeps the 32-bit resolution where it matters — near 0, where
        // the float is exact and the radius reaches sqrt(-2 ln 2^-33) = 6.76 sigma; near 1 it rounds to a multiple of
        // 2^-24 (radius steps of 3e-4 sigma around 0).  f32 logf / sqrtf (1 ulp) instead of the f64 pair of round 1:
        // the front-end kernel was 7 % of a BER batch, most of it here.
        float z[4];
        {
            const float r0 = sqrtf(-2.0f * logf(fminf(((float)r.x + 0.5f) * 2.3283064365386963e-10f, 1.0f)));
            const float r1 = sqrtf(-2.0f * logf(fminf(((float)r.z + 0.5f) * 2.3283064365386963e-10f, 1.0f)));
            float s0, c0, s1, c1;
            sincospif(2.0f * ((float)(r.y >> 8) + 0.5f) * 5.9604644775390625e-8f, &s0, &c0);
            sincospif(2.0f * ((float)(r.w >> 8) + 0.5f) * 5.9604644775390625e-8f, &s1, &c1);
            z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int t = t4 * 4 + i;
            if (t >= p.n_tx) break;
            uint32_t bit = tx_bit(t);
            float y = (bit ? 1.0f : -1.0f) + p.sigma * z[i];
            out[t] = p.llr_scale * y;
        }
    }
}

struct BackendParams {
    int k;
    uint32_t max_iter;
    uint64_t bch_max_errors;
    const uint8_t* decoded;      // [nframes][k] one 0/1 byte per info bit
    const int32_t* iters;        // [nframes]
    const uint32_t* messages;    // [nframes][ceil(k/32)]
    unsigned long long* counters;   // kBerCounters
};

__global__ void __launch_bounds__(256) ber_backend_kernel(BackendParams p) {
    const int kw = (p.k + 31) / 32;
    const uint8_t* dec = p.decoded + (size_t)blockIdx.x * p.k;
    const uint32_t* msg = p.messages + (size_t)blockIdx.x * kw;
    unsigned int errs = 0;
    for (int i = threadIdx.x; i < p.k; i += blockDim.x) errs += ((msg[i >> 5] >> (i & 31)) & 1u) != (uint32_t)dec[i];
    __shared__ unsigned int s_err;
    if (threadIdx.x == 0) s_err = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o; o >>= 1) errs += __shfl_xor_sync(0xffffffffu, errs, o);
    if ((threadIdx.x & 31) == 0 && errs) atomicAdd(&s_err, errs);
    __syncthreads();
    if (threadIdx.x == 0) {
        // ber.rs:313-337
        const unsigned long long be = s_err;
        const int32_t it = p.iters[blockIdx.x];
        const bool success = it >= 0;
        const unsigned long long iterations = success ? (unsigned long long)it : p.max_iter;
        atomicAdd(&p.counters[0], 1ull);
        if (be) { atomicAdd(&p.counters[1], be); atomicAdd(&p.counters[2], 1ull); }
        if (be && success) atomicAdd(&p.counters[3], 1ull);
        atomicAdd(&p.counters[4], iterations);
        if (!be) atomicAdd(&p.counters[5], iterations);
        if (be > p.bch_max_errors) { atomicAdd(&p.counters[6], be); atomicAdd(&p.counters[7], 1ull); }
        else atomicAdd(&p.counters[8], iterations);
    }
}

template <class T>
bool dev_upload(T** d, const std::vector<T>& h) {
    *d = nullptr;
    if (h.empty()) return true;
    if (cudaMalloc(d, h.size() * sizeof(T)) != cudaSuccess) { set_last_error("cudaMalloc failed (BER engine)"); return false; }
    return cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

}  // namespace

BerEngine::~BerEngine() {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    cudaFree(d_h0_ptr_); cudaFree(d_h0_idx_); cudaFree(d_g0_); cudaFree(d_kept_);
    for (Lane& ln : lanes_) {
        cudaFree(ln.d_llrs); cudaFree(ln.d_messages); cudaFree(ln.d_decoded); cudaFree(ln.d_iters); cudaFree(ln.d_counters);
        if (ln.h_counters) cudaFreeHost(ln.h_counters);
        if (ln.stream) cudaStreamDestroy(ln.stream);
        if (ln.done) cudaEventDestroy(ln.done);
    }
}

std::unique_ptr<BerEngine> BerEngine::create(const Graph& g, const DecoderImplementation& impl, const Puncturer* punct,
                                             const DecoderOptions& opt) {
    auto e = std::unique_ptr<BerEngine>(new BerEngine());
    std::string err;
    if (!EncoderPlan::from_graph(g, &e->plan_, &err)) { set_last_error(err); return nullptr; }
    e->decoder_ = build_decoder(impl, g, punct, opt);
    if (!e->decoder_) return nullptr;
    e->n_ = g.n; e->m_ = g.m; e->k_ = g.n - g.m;
    e->n_tx_ = g.n;
    std::vector<int> kept;
    if (punct) {
        if (!punct->puncture_map((size_t)g.n, &kept)) { set_last_error("codeword size not divisible by puncturing pattern length"); return nullptr; }
        e->n_tx_ = (int)kept.size();
    }
    // BerTest::new, ber.rs:246-259: N = round(n_cw / puncturer_rate), rate = k / N
    const double prate = punct ? punct->rate() : 1.0;
    const size_t n_frame = (size_t)llround((double)g.n / prate);
    e->rate_ = (double)e->k_ / (double)n_frame;
    if (cudaGetDevice(&e->device_) != cudaSuccess) return nullptr;
    if (e->plan_.staircase) {
        if (!dev_upload(&e->d_h0_ptr_, e->plan_.h0_ptr) || !dev_upload(&e->d_h0_idx_, e->plan_.h0_idx)) return nullptr;
    } else {
        // repack the 64-bit host rows as 32-bit words
        const int w32 = (e->k_ + 31) / 32;
        std::vector<uint32_t> g0((size_t)e->m_ * (size_t)w32, 0);
        for (int r = 0; r < e->m_; ++r)
            for (int w = 0; w < w32; ++w) {
                uint64_t v = e->plan_.g0[(size_t)r * (size_t)e->plan_.words + (size_t)(w >> 1)];
                g0[(size_t)r * (size_t)w32 + (size_t)w] = (uint32_t)(w & 1 ? v >> 32 : v);
            }
        e->g0_words_ = w32;
        if (!dev_upload(&e->d_g0_, g0)) return nullptr;
    }
    if (!dev_upload(&e->d_kept_, kept)) return nullptr;
    for (Lane& ln : e->lanes_) {
        if (cudaMalloc(&ln.d_counters, kBerCounters * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
        if (cudaMallocHost(&ln.h_counters, kBerCounters * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return e;
}

double BerEngine::noise_sigma(float ebn0_db) const {   // ber.rs:300-302; BITS_PER_SYMBOL modulation.rs:74, :151
    const double ebn0 = pow(10.0, 0.1 * (double)ebn0_db);
    const double esn0 = rate_ * (modulation_ == 1 ? 3.0 : 1.0) * ebn0;
    return sqrt(0.5 / esn0);
}

bool BerEngine::set_modulation(const std::string& name, int interleaving_columns) {
    int mod;
    if (name == "BPSK") mod = 0;                        // names: src/simulation/factory.rs:56-86
    else if (name == "8PSK") mod = 1;
    else { set_last_error("invalid modulation"); return false; }
    const int cols = interleaving_columns < 0 ? -interleaving_columns : interleaving_columns;
    // the reference panics on these at the first frame (modulation.rs:195, interleaving.rs:45)
    if (mod == 1 && n_tx_ % 3 != 0) { set_last_error("8PSK needs a frame length that is a multiple of 3 bits"); return false; }
    if (cols > 0 && n_tx_ % cols != 0) { set_last_error("frame length not divisible by the interleaver columns"); return false; }
    modulation_ = mod; il_cols_ = cols; il_backwards_ = interleaving_columns < 0 ? 1 : 0;
    return true;
}

bool BerEngine::ensure(Lane& ln, size_t nframes) {
    if (nframes <= ln.cap_frames) return true;
    cudaStreamSynchronize(ln.stream);
    cudaFree(ln.d_llrs); cudaFree(ln.d_messages); cudaFree(ln.d_decoded); cudaFree(ln.d_iters);
    ln.d_llrs = nullptr; ln.d_messages = nullptr; ln.d_decoded = nullptr; ln.d_iters = nullptr;
    const size_t kw = (size_t)(k_ + 31) / 32;
    if (cudaMalloc(&ln.d_llrs, nframes * (size_t)n_tx_ * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&ln.d_messages, nframes * kw * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&ln.d_decoded, std::max<size_t>(nframes * (size_t)k_, 1)) != cudaSuccess ||
        cudaMalloc(&ln.d_iters, nframes * sizeof(int32_t)) != cudaSuccess) {
        cudaGetLastError();
        set_last_error("cudaMalloc failed (BER engine buffers)");
        ln.cap_frames = 0;
        return false;
    }
    ln.cap_frames = nframes;
    return true;
}

// front-end -> decoder -> back-end -> counters to pinned host memory, all asynchronous on the lane's stream
bool BerEngine::enqueue(Lane& ln, int lane_index, float ebn0_db, uint32_t max_iterations, uint64_t first_frame, uint64_t nframes,
                        uint64_t seed, uint64_t bch_max_errors) {
    if (!ensure(ln, (size_t)nframes)) return false;
    const double sigma = noise_sigma(ebn0_db);
    FrontendParams fp{};
    fp.n = n_; fp.m = m_; fp.k = k_; fp.n_tx = n_tx_; fp.staircase = plan_.staircase ? 1 : 0;
    fp.h0_ptr = d_h0_ptr_; fp.h0_idx = d_h0_idx_; fp.g0 = d_g0_; fp.g0_words = g0_words_; fp.kept = d_kept_;
    fp.first_frame = first_frame;
    // key = (seed, Eb/N0 bits): a different noise/message stream for every operating point
    uint32_t eb_bits;
    memcpy(&eb_bits, &ebn0_db, 4);
    fp.seed_lo = (uint32_t)seed; fp.seed_hi = (uint32_t)(seed >> 32) ^ eb_bits;
    fp.sigma = (float)sigma; fp.llr_scale = (float)(-2.0 / (sigma * sigma));
    fp.modulation = modulation_; fp.il_cols = il_cols_; fp.il_backwards = il_backwards_; fp.sigma_d = sigma;
    fp.llrs = ln.d_llrs; fp.messages = ln.d_messages;
    const size_t smem = ((size_t)(k_ + 31) / 32 + (size_t)(m_ + 31) / 32 + 4) * sizeof(uint32_t);
    if (smem > 48 * 1024) LDPC_CUDA_CHECK(cudaFuncSetAttribute(ber_frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LDPC_CUDA_CHECK(cudaMemsetAsync(ln.d_counters, 0, kBerCounters * sizeof(unsigned long long), ln.stream));
    ber_frontend_kernel<<<(unsigned)nframes, 256, smem, ln.stream>>>(fp);
    LDPC_CUDA_CHECK(cudaGetLastError());
    if (!decoder_->decode_batch_device_lane(lane_index, ln.d_llrs, false, (size_t)n_tx_, (size_t)nframes, max_iterations, ln.d_decoded, (size_t)k_,
                                            (size_t)k_, ln.d_iters, ln.stream))
        return false;
    BackendParams bp{};
    bp.k = k_; bp.max_iter = max_iterations; bp.bch_max_errors = bch_max_errors;
    bp.decoded = ln.d_decoded; bp.iters = ln.d_iters; bp.messages = ln.d_messages; bp.counters = ln.d_counters;
    ber_backend_kernel<<<(unsigned)nframes, 256, 0, ln.stream>>>(bp);
    LDPC_CUDA_CHECK(cudaGetLastError());
    LDPC_CUDA_CHECK(cudaMemcpyAsync(ln.h_counters, ln.d_counters, kBerCounters * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ln.stream));
    launches_ += 2 + 3;
    return true;
}

int64_t BerEngine::submit(float ebn0_db, uint32_t max_iterations, uint64_t first_frame, uint64_t nframes, uint64_t seed,
                          uint64_t bch_max_errors) {
    if (cudaSetDevice(device_) != cudaSuccess) { set_last_error("cudaSetDevice failed"); return -1; }
    const int64_t ticket = next_ticket_;
    Lane& ln = lanes_[ticket & 1];
    if (ln.ticket >= 0) { set_last_error("BER engine: two batches are already in flight (wait for the older ticket first)"); return -1; }
    if (nframes > 0) {
        if (!enqueue(ln, (int)(ticket & 1), ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors)) return -1;
    } else {
        memset(ln.h_counters, 0, kBerCounters * sizeof(unsigned long long));
    }
    if (cudaEventRecord(ln.done, ln.stream) != cudaSuccess) { set_last_error("cudaEventRecord failed"); return -1; }
    ln.ticket = ticket;
    ++next_ticket_;
    return ticket;
}

bool BerEngine::wait(int64_t ticket, uint64_t* counters) {
    if (ticket < 0) { set_last_error("BER engine: bad ticket"); return false; }
    Lane& ln = lanes_[ticket & 1];
    if (ln.ticket != ticket) { set_last_error("BER engine: ticket is not in flight"); return false; }
    ln.ticket = -1;
    LDPC_CUDA_CHECK(cudaSetDevice(device_));
    LDPC_CUDA_CHECK(cudaEventSynchronize(ln.done));
    for (int i = 0; i < kBerCounters; ++i) counters[i] += ln.h_counters[i];
    return true;
}

bool BerEngine::run(float ebn0_db, uint32_t max_iterations, uint64_t first_frame, uint64_t nframes, uint64_t seed,
                    uint64_t bch_max_errors, uint64_t* counters, float* dump_llrs, uint8_t* dump_decoded, int32_t* dump_iters,
                    uint32_t* dump_messages) {
    if (nframes == 0) return true;
    const int64_t t = submit(ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors);
    if (t < 0) return false;
    Lane& ln = lanes_[t & 1];
    // test hooks: the lane's buffers are intact until its next submit.  A failed copy must not leave the ticket in
    // flight, so the lane is always waited for.
    auto dump = [&](void* dst, const void* src, size_t bytes) {
        return !dst || cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ln.stream) == cudaSuccess;
    };
    const bool dumped = dump(dump_llrs, ln.d_llrs, nframes * (size_t)n_tx_ * sizeof(float)) &&
                        dump(dump_decoded, ln.d_decoded, nframes * (size_t)k_) &&
                        dump(dump_iters, ln.d_iters, nframes * sizeof(int32_t)) &&
                        dump(dump_messages, ln.d_messages, nframes * ((size_t)(k_ + 31) / 32) * sizeof(uint32_t)) &&
                        cudaStreamSynchronize(ln.stream) == cudaSuccess;
    if (!dumped) {
        set_last_error(std::string("BER engine: dump copy failed: ") + cudaGetErrorString(cudaGetLastError()));
        uint64_t scratch[kBerCounters] = {};
        wait(t, scratch);
        return false;
    }
    return wait(t, counters);
}

}  // namespace ldpc
