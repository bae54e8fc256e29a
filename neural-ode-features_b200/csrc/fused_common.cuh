// Shared declarations of the fused ODE-Net route: workspace layout, launch arguments, geometry.
#pragma once
#include <cstdlib>
#include "node_common.cuh"
#include "ptx.cuh"

namespace node {

constexpr int kC = 64;                 // channels of the fused kernels (n_filters=64)
constexpr int kGroups = 32;            // GroupNorm(min(32, C), C)
constexpr int kCpg = kC / kGroups;     // channels per group
constexpr int kFThreads = 256;

constexpr int kWTileBytes = 2 * 2 * 64 * 128;  // hi/lo x kblock x 64 rows x 128 B = 32 KB per tap
constexpr int kTmemCols = 512;
constexpr int kAccCols = 128;          // up to two 128-row accumulators of 64 columns
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr int kPartialBlocksF = 296;
constexpr int kMaxGrid = 148;

enum { MODE_F0 = 0, MODE_PROBE = 1, MODE_STEP = 2, MODE_EVAL = 3 };
// convolution engines (C ABI `conv_mode`): the fp32 contract is met by the split engines (x3)
enum { CONV_F16X3 = 0, CONV_F16 = 1, CONV_SIMT = 2, CONV_TF32X3 = 3, CONV_TF32 = 4 };

constexpr int kW16TileBytes = 128 * 128;       // f16 engine: 128 rows (64 hi + 64 lo cout) x 64 cin halves, SW128
constexpr int kW16Sets = 4;                    // conv1, conv2, and their transposed/flipped twins (adjoint dgrad)

struct FusedWs {
  node_ctl_t* ctl; double* sums; int* nonfinite; double* partials; double* t_out;
  float* wtiles;  // [2 conv][9 tap][hi/lo][2 kblock][64 cout][32 cin] swizzled
  float* wraw;    // [2][C][C+1][9] copies of the live weights (SIMT engine)
  float* tmap;    // [2][C][HW]
  float* bias;    // [2][C]
  float* gn;      // [3][2][C] gamma, beta
  uint16_t* w16;  // f16 engine: [kW16Sets][9 tap][128 rows][64 halves], UMMA K-major SWIZZLE_128B image
  uint16_t* w16p; // CTA-pair engine (step8_engine.cuh): [2 conv][9 tap][2 cta][64 rows][64 halves] half tiles, SWIZZLE_128B image
  uint16_t* w16q; // adjoint pair engine (vjp8_engine.cuh): [4 sets][9 tap][2 cta][32 hi rows + 32 lo rows][64 halves], SWIZZLE_128B image
  float* tmapc;   // [2][9 border classes][C]: the 9 distinct values of Tmap per channel
  float* scal;    // [16] power-of-two operand scales of the f16 engine (see odefunc_step.cu)
  float* Y[2]; float* F[2]; float* K[5]; float* YMID;
};

struct Geo { int N, H, W, HW, G, MT, ngroups, gs; };    // gs: images per super-tile the strip engines use for this batch (strip_gs)

struct FusedArgs {
  FusedWs w; Geo g;
  int mode, conv_mode, nw;
  const float* y_in; float* k_out; float* out0;
  float t_explicit, tsign, eps;
};

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
constexpr int kMaxT = 1024;

static int64_t ws_layout(void* base, int N, int C, int H, int W, FusedWs* out) {
  const int64_t E = (int64_t)N * C * H * W;
  const int64_t HW = (int64_t)H * W;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o = align_up(o + bytes, 1024); return r; };
  const int64_t o_ctl = take(sizeof(node_ctl_t));
  const int64_t o_sums = take(sizeof(double) * 2 * NODE_MAX_SEG);
  const int64_t o_nf = take(sizeof(int));
  const int64_t o_part = take(sizeof(double) * 2 * NODE_MAX_SEG * kPartialBlocksF);
  const int64_t o_tout = take(sizeof(double) * kMaxT);
  const int64_t o_wt = take((int64_t)2 * 9 * kWTileBytes);
  const int64_t o_wraw = take((int64_t)2 * C * (C + 1) * 9 * 4);
  const int64_t o_tmap = take((int64_t)2 * C * HW * 4);
  const int64_t o_bias = take((int64_t)2 * C * 4);
  const int64_t o_gn = take((int64_t)6 * C * 4);
  const int64_t o_w16 = take((int64_t)kW16Sets * 9 * kW16TileBytes);
  const int64_t o_w16p = take((int64_t)2 * 9 * kW16TileBytes);
  const int64_t o_w16q = take((int64_t)4 * 9 * kW16TileBytes);
  const int64_t o_tmapc = take((int64_t)2 * 9 * C * 4);
  const int64_t o_scal = take(16 * 4);
  int64_t o_state[10];
  for (int i = 0; i < 10; ++i) o_state[i] = take(E * 4);
  if (out != nullptr) {
    char* b = (char*)base;
    out->ctl = (node_ctl_t*)(b + o_ctl); out->sums = (double*)(b + o_sums); out->nonfinite = (int*)(b + o_nf);
    out->partials = (double*)(b + o_part); out->t_out = (double*)(b + o_tout);
    out->wtiles = (float*)(b + o_wt); out->wraw = (float*)(b + o_wraw); out->tmap = (float*)(b + o_tmap);
    out->bias = (float*)(b + o_bias); out->gn = (float*)(b + o_gn);
    out->w16 = (uint16_t*)(b + o_w16); out->w16p = (uint16_t*)(b + o_w16p); out->w16q = (uint16_t*)(b + o_w16q); out->tmapc = (float*)(b + o_tmapc); out->scal = (float*)(b + o_scal);
    out->Y[0] = (float*)(b + o_state[0]); out->Y[1] = (float*)(b + o_state[1]);
    out->F[0] = (float*)(b + o_state[2]); out->F[1] = (float*)(b + o_state[3]);
    for (int i = 0; i < 5; ++i) out->K[i] = (float*)(b + o_state[4 + i]);
    out->YMID = (float*)(b + o_state[9]);
  }
  return o;
}

static int strip_gs(int N, int H, int W);
static bool make_geo(int N, int C, int H, int W, Geo* g) {
  if (C != kC || N < 1 || H < 1 || W < 1) return false;
  const int HW = H * W;
  if (HW > 256) return false;
  g->N = N; g->H = H; g->W = W; g->HW = HW;
  g->G = HW <= 128 ? 128 / HW : 1;
  g->MT = (g->G * HW + 127) / 128;
  g->ngroups = (N + g->G - 1) / g->G;
  g->gs = strip_gs(N, H, W);
  return true;
}

// ---- adjoint kernels (K7): launch arguments shared by vjp_engine.cuh, wgrad_engine.cuh and odefunc_vjp.cu
constexpr int kWgCols = 80;                      // weight-gradient accumulator columns per tap: 64 ci + ones + 15 unused
constexpr int kWgSplits = 37;                    // 37 splits x 2 tap groups x 2 convolutions = 148 CTAs

struct VjpArgs {
  FusedWs w; Geo g;
  const float* y; const float* adj;
  float* f_out; float* vy_out;
  float* R[2]; float* GC[2];     // [N,64,H,W] fp32, operands of the weight-gradient GEMM
  float* chan_part;              // [grid][6][64]: dgamma1, dbeta1, dgamma2, dbeta2, dgamma3, dbeta3
  double* t_part;                // [grid]: sum GC*Tmap over both convolutions
  const float* t_dev;            // device scalar: the time the solver evaluates at
  unsigned* gc_max;              // [2] bit patterns of max |GC1|, |GC2| over the batch (zeroed before the launch): k_wgrad's scale
  float tsign, eps;
};

// CTAs of a weight-gradient launch = splits x 2 tap groups x pairs: whole waves of the 148 SMs (1 CTA per SM)
static inline int wgrad_splits(int NST, int ncv) {
  const int waves = (ncv + 2) / 3;
  int n = 148 * waves / (2 * ncv);
  if (ncv == 2) n = kWgSplits;
  return NST < n ? NST : n;
}
// the dense 8x8 weight-gradient kernel (wgrad8_engine.cuh): stages of two images
static inline int wgrad8_splits(int N, int ncv) {
  const int NST = (N + 1) / 2, waves = (ncv + 2) / 3;
  const int n = 148 * waves / (2 * ncv);
  return NST < n ? NST : n;
}
constexpr int kWgMaxPairs = 6;                   // (input, output-gradient) pairs one k_wgrad launch serves (grid.z)
struct WgradArgs {
  Geo g;
  const float* R[kWgMaxPairs]; const float* GC[kWgMaxPairs];   // [N,64,H,W] fp32 (the adjoint's pairs are written by k_vjp)
  float* part;                                   // [splits][ncv][9 tap][64 co][kWgCols]
  const unsigned* gc_max[kWgMaxPairs];           // per pair: bit pattern of max |GC| over the batch (k_vjp / k_absmax)
  const float* scal[kWgMaxPairs];                // per pair: the power-of-two scale of the input operand (device scalar)
  int nsplit, ncv;
};

// images per super-tile of the position-strip tiling (Tile<H,W>::G in step_engine.cuh)
__host__ __device__ constexpr int strip_images(int H, int W) {
  const int Wp = W + 1, IS = (H + 1) * Wp, SPAN = (H - 1) * Wp + W;
  const int MT = SPAN <= 256 ? 2 : (SPAN + 127) / 128;
  return (MT * 128 - SPAN) / IS + 1;
}

// Images per super-tile actually used by the strip engines (k_step, k_vjp): a super-tile can hold strip_images() of them, but a
// small batch is spread over the SMs first - at the reference's batch of 128 every CTA gets ONE image and, with the M tiles that
// hold no image skipped by the issuer, a convolution job is half the MMAs (the evaluation is a dependent chain: 30 -> 25 us).
// NODE_B200_STRIP_GS=0 keeps full super-tiles.
static int strip_gs(int N, int H, int W) {
  const int G = strip_images(H, W);
  static const char* e = getenv("NODE_B200_STRIP_GS");
  if (e != nullptr && e[0] == '0') return G;
  const int gs = (N + kMaxGrid - 1) / kMaxGrid;
  return gs < 1 ? 1 : (gs > G ? G : gs);
}

// f16 engine (odefunc_step.cu)
bool step_engine_supports(int H, int W);
int launch_step_engine(const FusedArgs& a, cudaStream_t st);
int launch_prepare16(const FusedWs& w, int H, int W, const float* c1w, const float* c2w, const float* g1w, const float* g1b,
                     const float* g2w, const float* g2b, cudaStream_t st);
int launch_prepare_dgrad(const FusedWs& w, const float* c1w, const float* c2w, cudaStream_t st);   // odefunc_vjp.cu

}  // namespace node
