/* qtb200.h -- C ABI of the B200-native QuantTorch quantized-forward kernels.
 *
 * One shared library (libqtb200.so, built by __graft_entry__.build() with
 * nvcc -gencode arch=compute_100a,code=sm_100a).  Plain C: device pointers,
 * sizes, POD structs, a cudaStream_t passed as void*.  No torch types.
 *
 * The reference (Enderdead/Pytorch_Quantize_impls, "QuantTorch") has no FFI of
 * its own: every quantized layer is a Python fake-quant op followed by a dense
 * fp32 torch.nn.functional.linear / conv2d.  Each entry point below therefore
 * cites the reference Python code (path:line under /root/reference) whose work
 * it replaces; the Python binding a maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a negative QT_E* code on failure;
 *     qt_last_error() returns a thread-local message for the last failure.
 *   - all data pointers are DEVICE pointers owned by the caller; nothing is
 *     allocated, freed or synchronised inside the library; every launch goes
 *     to the stream passed in (stream-ordered, re-entrant, thread-safe).
 *   - matrices are row-major; "ld" arguments are in ELEMENTS of that matrix.
 *   - bit-packed rows: bit i of uint32 word j <-> column 32 j + i; 1 <-> +1.
 */
#ifndef QTB200_H
#define QTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QT_VERSION 105

enum {
  QT_OK = 0,
  QT_EINVAL = -1,      /* bad argument (null pointer, misaligned ld, unsupported bit width ...) */
  QT_ECUDA = -2,       /* CUDA runtime / driver error (message holds cudaGetErrorString) */
  QT_EUNSUPPORTED = -3 /* device is not sm_100 or shape not supported by the requested kernel */
};

int qt_version(void);
/* sizeof() of a struct of this header by name ("QtEpilogue", ...), -1 if unknown: lets a binding check its mirror. */
int qt_sizeof(const char* struct_name);
const char* qt_last_error(void);
/* sm major/minor, SM count, 1 if tcgen05 kernels can run on `device`. */
int qt_device_caps(int device, int* sm_major, int* sm_minor, int* num_sms, int* has_tcgen05);

/* ------------------------------------------------------------------------
 * Activation quantizers (elementwise + per-row reductions), one pass over x.
 * Replaces the 3..10 separate ATen elementwise passes of
 *   safeSign                     QuantTorch/functions/common.py:4-7
 *   BinaryConnectDeterministic   QuantTorch/functions/binary_connect.py:22-28
 *   TernaryConnectDeterministic  QuantTorch/functions/terner_connect.py:24-27
 *   _quantize / DorefaQuant      QuantTorch/functions/dorefa_connect.py:11-25,49-63
 *   _quantOpXnor (dim=1)         QuantTorch/functions/xnor_connect.py:20-28
 *   LogQuant / LinQuant          QuantTorch/functions/log_lin_connect.py:29-32,61-68
 * and additionally emits the low-bit operand the contraction kernels consume.
 * ---------------------------------------------------------------------- */
enum {
  QT_Q_SIGN = 0,     /* +1 iff !(x<0)                         codes {-1,+1}          */
  QT_Q_TERNARY = 1,  /* x>=.5 -> 1, -.5<=x<.5 -> 0, else -1   codes {-1,0,+1}        */
  QT_Q_DOREFA = 2,   /* c = rint((2^k-1) x), y = fl(1/n) c    codes c (unclamped)    */
  QT_Q_XNOR_ROW = 3, /* sign(x) * mean(x,row)  (torch.sign)   codes {-1,0,+1}, row_scale = mean */
  QT_Q_LOG = 4,      /* sign(x) 2^clamp(round(log2|x|), fsr-2^bw, fsr)   codes = the value (bf16 lanes)       */
  QT_Q_LIN = 5,      /* sign(x) clamp(round(|x|/step) step, 0, 2^fsr)    codes = value / step (int8 / uint8 / bf16 lanes) */
  QT_Q_SPLIT = 6     /* no quantisation: bf16 hi/lo (codes_kind 4) or hi/mid/lo (codes_kind 6) split of x */
};

typedef struct QtActQuant {
  int mode;             /* QT_Q_* */
  int bit_width;        /* DoReFa k (2..8), Log/Lin bit width */
  int fsr;              /* Log/Lin full-scale range */
  int with_sign;        /* Log/Lin */
  const float* x;       /* [rows, ld_x] fp32 */
  int64_t rows, cols, ld_x;
  float* y;             /* optional fp32 fake-quant result [rows, ld_y] (the reference op's return value) */
  int64_t ld_y;
  void* codes;          /* optional low-bit operand, [rows, ld_codes] of int8/uint8 or bf16 (see codes_kind);
                           columns cols..ld_codes-1 are zero-filled */
  int codes_kind;       /* 0 = none, 1 = int8, 2 = uint8, 3 = bf16, 4 = bf16 hi/lo planes (plane stride = rows*ld_codes),
                           5 = fp16, 6 = bf16 hi/mid/lo planes (24 significant bits: the fp32-faithful split),
                           7 = fp4 (e2m1) codes, two per byte, element 2j in the low nibble of byte j; integer codes in
                               [-4, 4] only (QT_Q_SIGN, QT_Q_TERNARY, QT_Q_DOREFA k = 2); ld_codes % 32 == 0 */
  int64_t ld_codes;     /* in ELEMENTS (codes_kind 7: two elements per byte) */
  uint32_t* bits;       /* optional bit-packed sign rows [rows, ld_bits] (QT_Q_SIGN only) */
  int64_t ld_bits;
  int32_t* row_sum;     /* optional [rows]: sum over columns of the integer codes */
  float* row_scale;     /* optional [rows]: QT_Q_XNOR_ROW row mean */
  int32_t* overflow;    /* optional device flag, OR-ed with 1 when a code does not fit the int8/uint8 lane */
  /* optional fused pre-transform (inference-time BatchNorm + activation clamp in front of the quantizer,
     e.g. benchmark/BinaryNet/AlexNetBin.py:14-16 `BatchNorm2d -> Hardtanh -> BinaryConnect`):
       x' = clamp(x * pre_scale[ch] + pre_shift[ch], pre_lo, pre_hi),   ch = (column / pre_hw) % pre_channels
     pre_scale == NULL disables it; pre_clamp == 0 skips the clamp. */
  const float* pre_scale;
  const float* pre_shift;
  int64_t pre_channels, pre_hw;
  int pre_clamp;
  float pre_lo, pre_hi;
  int64_t nhwc_c;       /* 0: codes are row-major [rows, ld_codes].  C > 0: x (and y) are NCHW with rows = B images of
                           C channels x (cols / C) pixels, and the int8/uint8 codes are written channels-last
                           [B, H*W, C] (dense) -- the layout the conv gather / TMA im2col reads with 16-byte vectors */
  int row_parts;        /* QT_Q_XNOR_ROW: capacity of row_scale in [rows]-sized parts (0 or 1: one vector holding the mean).
                           With P = qt_quant_xnor_parts(cols, y != NULL, row_parts) > 1 the rows are processed in P column
                           chunks and row_scale[p * rows + r] receives the partial SUM of chunk p (mean = sum_p / cols; the
                           contraction epilogue adds them in order: QtEpilogue.row_scale_parts / row_scale_mul) */
  int max_ctas;         /* 0: one CTA per 8 row chunks (fastest when the quantizer has the GPU to itself).  > 0: at most this many
                           CTAs, each looping over row chunks -- a bounded footprint (8 warps, ~8 K registers, no shared memory
                           per CTA) that shares every SM with a persistent tcgen05 contraction running on another stream
                           (the quantizer of row band i+1 beside the product of band i).  Code-only calls only (no y / bits) */
  int32_t* ready;       /* optional progress counters (device, zeroed by the caller): after the codes of a 1024-column chunk of row r
                           are stored (and fenced), ready[r / ready_rows] is incremented -- a consumer kernel running CONCURRENTLY
                           (QtEpilogue.a_ready) starts on a row block as soon as its count reaches ready_rows * cols / 1024.
                           Lean code-only calls only (cols % 1024 == 0); ignored otherwise is an error */
  int ready_rows;
} QtActQuant;

int qt_quant_act(const QtActQuant* p, void* stream);
/* Number of row_scale parts qt_quant_act writes for a QT_Q_XNOR_ROW call (1 = the mean itself). */
int qt_quant_xnor_parts(int64_t cols, int has_y, int capacity);

/* ------------------------------------------------------------------------
 * Weight quantizers / packers: fp32 master weights [n, k] -> k-bit HBM format.
 *   BinaryNet  sign bits                     binary_layers.py:42-46 (bin_op on W)
 *   Terner     two bit planes (nz, sign)     terner_layers.py:47-51
 *   DoReFa     k-bit codes c in [0, 2^k-1]   dorefa_connect.py:99-111 (nnQuantWeight)
 *              k == 1: sign bits + scalar E = mean|W|
 *   XnorNet    sign/nz planes + alpha[k] = mean(|W|, dim 0)   xnor_connect.py:111-112
 * `stats` is a small device scratch (>= 16 floats) that receives the reductions:
 *   stats[0] = max|tanh W| (DoReFa k>=2)  stats[1] = mean|W| (DoReFa k==1)
 *   stats[2] = number of exact zeros (as float)   stats[3] = max|W|
 * ---------------------------------------------------------------------- */
enum {
  QT_W_SIGN = 0,     /* packed: bits[n, ldw]                                  */
  QT_W_TERNARY = 1,  /* packed: nz[n, ldw] then sign[n, ldw] (plane stride n*ldw words) */
  QT_W_DOREFA = 2,   /* packed: codes, lane bits = 1,2,4,8 (k=3 -> 4, k=5..7 -> 8), little-endian within a byte */
  QT_W_XNOR = 3      /* packed: nz/sign planes as ternary (torch.sign keeps zeros) + alpha[k] */
};

typedef struct QtWeightPack {
  int mode;          /* QT_W_* */
  int bit_width;     /* DoReFa k */
  const float* w;    /* [n, ld_w] fp32 */
  int64_t n, k, ld_w;
  void* packed;      /* see mode */
  int64_t ld_packed; /* row stride of `packed` in BYTES (multiple of 4) */
  float* alpha;      /* QT_W_XNOR: [k]; written unless alpha_is_input */
  int alpha_is_input; /* 1: alpha was computed by the caller (XNORConv2d per-tap alpha, xnor_connect.py:140) */
  float* stats;      /* device scratch, >= 16 floats */
  float* wq;         /* optional fp32 fake-quant weights [n, ld_w] (what the reference's weight op returns) */
} QtWeightPack;

int qt_pack_weight(const QtWeightPack* p, void* stream);

/* alpha[c] = mean over the n rows of |w[r, c]|  (torch.mean(torch.abs(W), 0), xnor_connect.py:111). */
int qt_col_absmean(const float* w, int64_t n, int64_t k, int64_t ld_w, float* alpha, void* stream);

/* Expand a packed weight matrix into the transient operand the tensor-core
 * kernels read (written and re-read through L2; never the persistent format).
 *   kind 1: int8  centred codes  (sign: +-1, ternary: -1/0/1, DoReFa k<=7: 2c-n)
 *   kind 2: uint8 raw codes c    (DoReFa k == 8; zero point handled in the epilogue)
 *   kind 3: bf16 exact values of kind 1
 *   kind 4: bf16 hi/lo planes of alpha[k] * sign (XnorNet; plane stride = n*ld_out elements)
 *   kind 5: fp16 alpha[k] * sign, one plane (XnorNet fast route; pass alpha pre-normalised to max 1 so that the
 *           values sit in fp16's normal range, and put the max back through the epilogue's col_scale)
 *   kind 6: fp16 exact values of kind 1 (integers up to 255 are exact in fp16)
 *   kind 7: fp4 (e2m1) centred codes, two per byte (sign: +-1, ternary: -1/0/1, DoReFa k<=2: 2c-n in {-3,-1,1,3});
 *           ld_out % 32 == 0.  The operand of qt_gemm_f4.
 */
typedef struct QtWeightExpand {
  int mode, bit_width;
  const void* packed;
  int64_t n, k, ld_packed;
  const float* alpha;   /* QT_W_XNOR */
  void* out;
  int out_kind;
  int64_t ld_out;       /* elements; columns k..ld_out-1 are zero-filled */
} QtWeightExpand;

int qt_expand_weight(const QtWeightExpand* p, void* stream);

/* LogLin layers (log_lin_layers.py:6-93, log_lin_connect.py:9-80) keep their weights as int8 codes in HBM:
 *   lin: code = sign * round(|w| / step) in [-2^bw, 2^bw], value = code * step, step = 2^(fsr - bw)
 *        -> with Lin-quantized activations (QT_Q_LIN, int8 codes) the layer is an exact qt_gemm_i8 with scale step_a * step_w;
 *   log: code = sign * (e - emin + 1), 0 <-> 0, value = sign * 2^e, emin = fsr - 2^bw.
 * qt_expand_loglin writes the bf16 operand of qt_gemm_f16 (lin: the integer code, log: the power of two; both exact). */
int qt_expand_loglin(const void* codes, int64_t n, int64_t k, int64_t ld_codes, int is_log, int emin, void* out, int64_t ld_out,
                     void* stream);

/* ------------------------------------------------------------------------
 * im2col gather for the conv layers (binary_layers.py:103-106, terner_layers.py:89-92,
 * dorefa_layers.py:77-82, xnor_connect.py:139-146).  Input NCHW (nhwc = 0) or channels-last NHWC
 * (nhwc = 1, 16-byte vector copies when C/groups * elem_bytes % 16 == 0), zero padding
 * (a padded tap contributes exactly 0, as F.conv2d does).  Output row m = (b, oh, ow),
 * column = (kh, kw, c) of group `g` -- channel fastest, so weights are packed from
 * W.permute(0, 2, 3, 1); element type = 1, 2 or 4 bytes (int8/uint8 codes, bf16, fp32).
 * Optional row_sum of int8/uint8 codes.
 * ---------------------------------------------------------------------- */
typedef struct QtIm2col {
  const void* x;     /* [B, C, H, W] or [B, H, W, C] */
  int nhwc;          /* input layout */
  int elem_bytes;    /* 1, 2, 4 */
  int is_unsigned;   /* for row_sum of 1-byte codes */
  int64_t B, C, H, W;
  int kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int groups, group;           /* channels [group*C/groups, (group+1)*C/groups) */
  int64_t OH, OW;
  void* out;                   /* [B*OH*OW, ld_out] */
  int64_t ld_out;              /* elements; columns beyond (C/groups)*kh*kw zero-filled */
  int32_t* row_sum;            /* optional */
  int split3;                  /* != 0: x is fp32 (elem_bytes 4) and `out` receives bf16 planes of the gathered values, plane
                                  stride = B*OH*OW*ld_out elements (fused gather + split for first layers that see real-valued
                                  images): 1 or 3 = THREE planes hi/mid/lo (24 significant bits, fp32-faithful), 2 = TWO planes
                                  hi/lo (16 significant bits, relative error <= 2^-17) */
} QtIm2col;

int qt_im2col(const QtIm2col* p, void* stream);

/* ------------------------------------------------------------------------
 * Contractions.  All compute  D[m, n] = sum_k A[m, k] * W[n, k]  (i.e. x . W^T,
 * the F.linear call at binary_layers.py:44, terner_layers.py:49, dorefa_layers.py:43,
 * xnor_connect.py:115) and apply the fused epilogue
 *     t = acc_mul * acc + rs_mul * row_sum[m]                (integer kernels; exact)
 *     y = float(t) * scale * row_scale[m] * col_scale[n] + bias[n]
 * Output addressing:
 *     out_mode 0 (row major) : out[m * ldo + n]
 *     out_mode 1 (NCHW)      : out[((m / P) * ldo + n) * P + m % P],  P = nchw_inner (= OH*OW), ldo = total
 *                              output channels (a conv group writes at out + group_offset * P)
 * acc_out (optional, integer kernels) receives the raw int32 accumulators [M, N] row major.
 * ---------------------------------------------------------------------- */
/* Optional fused re-quantisation of the layer output (the inter-layer pattern every reference net repeats:
 *   quantized linear/conv -> [BatchNorm (eval)] -> [Hardtanh / ReLU] -> activation quantizer -> next quantized layer,
 * benchmark/BinaryNet/MLPBin.py:42-53, AlexNetBin.py:13-48, models/samples/AlexNet_Dorefa.py:46-84).  The tcgen05 epilogue
 * applies  y' = clamp(y, lo, hi)  (the caller folds the BatchNorm affine into col_scale / bias) and the quantizer `mode`
 * to the fp32 value y it just produced, and writes the NEXT layer's low-bit operand instead of (or besides) the fp32
 * tensor: the hidden activation never exists in fp32 in HBM.
 *   codes[m, n]  row-major [M, ld_codes]: for a linear layer that is the next layer's [rows, K] operand, for a conv layer
 *                (m = (b, oh, ow), n = output channel) it is the channels-last [B, OH, OW, C] code tensor qt_conv_i8 reads.
 *   QT_Q_XNOR_ROW needs the mean of y' over the whole row, which no single tile sees: the epilogue writes partial row
 *   sums  row_part[p * M + m]  (p = index of the 32-column chunk group that produced it, row_parts of them in total; the
 *   library stores the count in row_parts) and the consumer sums them in a fixed order (QtEpilogue.row_scale_parts /
 *   row_scale_mul = 1/N): deterministic, no atomics.  row_sum_part does the same for the integer code sums the unsigned
 *   DoReFa-8 weight zero point needs.
 * Only the tcgen05 kernels implement it; other routes return QT_EUNSUPPORTED when `requant` is set. */
typedef struct QtRequant {
  int mode;             /* QT_Q_SIGN, QT_Q_TERNARY, QT_Q_DOREFA or QT_Q_XNOR_ROW */
  int bit_width;        /* DoReFa k (2..8) */
  void* codes;          /* [M, ld_codes] */
  int codes_kind;       /* 1 int8, 2 uint8, 3 bf16, 5 fp16, 7 fp4 (e2m1, two per byte) -- as QtActQuant.codes_kind */
  int64_t ld_codes;     /* ELEMENTS; multiple of 32 and >= N rounded up to 32; columns N..round_up(N,32)-1 are zero-filled */
  int clamp;            /* 1: y' = min(max(y, lo), hi) before the quantizer */
  float lo, hi;
  float* row_part;      /* QT_Q_XNOR_ROW: [>= qt_requant_max_parts(N), M] partial sums of y' */
  int32_t* row_sum_part;/* optional: [>= qt_requant_max_parts(N), M] partial sums of the integer codes */
  int row_parts;        /* OUT: number of partial rows written */
  int32_t* overflow;    /* optional sticky device flag (a DoReFa code left its lane) */
  int64_t cover;        /* 0: columns N .. ld_codes-1 of a row are zero-filled (channel padding).  > 0: only columns N .. cover-1
                           are (the row pitch is wider than this launch's slice: two output-column phases of a W-folded conv
                           interleave their pixels, ld_codes = 2 * channel pitch) */
} QtRequant;

/* upper bound on QtRequant.row_parts for an N-column output */
int qt_requant_max_parts(int64_t N);

typedef struct QtEpilogue {
  const float* bias;      /* [N] or NULL */
  const float* row_scale; /* [M] or NULL */
  const float* col_scale; /* [N] or NULL */
  const int32_t* row_sum; /* [M] or NULL */
  float scale;
  int32_t acc_mul, rs_mul;
  float* out;
  int64_t ldo;
  int out_mode;
  int64_t nchw_inner;
  int32_t* acc_out;
  /* ---- fused inter-layer chain (all optional; zero-initialise when unused) ---- */
  QtRequant* requant;     /* NULL: none.  With a requant, `out` may be NULL (code-only chain) */
  int row_scale_parts;    /* > 0: row_scale is [row_scale_parts, M] partial sums written by a previous layer's requant;
                             the effective row scale is row_scale_mul * sum_p row_scale[p * M + m] */
  float row_scale_mul;
  int row_sum_parts;      /* > 0: row_sum is [row_sum_parts, M] partial sums, summed the same way */
  int out_clamp;          /* 1: y = min(max(y, out_lo), out_hi) before it is written (a Hardtanh / ReLU / ReLU6 that follows the
                             layer, e.g. models/Resnet/Resnet_bin.py:27-31, costs no pass over the activation) */
  float out_lo, out_hi;
  const float* residual;  /* optional fp32 [M, ld_res] (row major): y += residual[m, n] after scale / bias and BEFORE out_clamp and the
                             requant -- the shortcut add of a residual block (models/Resnet/Resnet_bin.py:27-31
                             `out += self.shortcut(x); out = F.relu(out)`) folded into the second conv's epilogue, with the
                             channels-last fp32 activation as [pixels, channels] matrix.  out_mode 0 only */
  int64_t ld_res;
  /* Operand dependency (tcgen05 GEMM routes, plain row-major A): the A rows are being produced by a kernel running on another
     stream at the same time (qt_quant_act with `ready` counters).  Before the TMA producer loads rows of block
     b = row / a_ready_rows it waits until a_ready[b] >= a_ready_target (acquire at gpu scope + proxy fence).  The quantizer of
     a layer then runs BESIDE its contraction instead of in front of it (HBM-bound pass hidden behind the tensor-bound one)
     without cutting the product into bands.  a_ready_rows must be a multiple of 128.  NULL: no dependency. */
  const int32_t* a_ready;
  int a_ready_rows, a_ready_target;
} QtEpilogue;

/* 1-bit x 1-bit: acc = K - 2 popc(a ^ w).  CUDA-core XNOR + popcount. */
int qt_gemm_b1b1(const uint32_t* a_bits, int64_t lda_words, const uint32_t* w_bits, int64_t ldw_words,
                 int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, void* stream);

/* 1-bit activations x ternary weights (two planes): acc = popc(nz) - 2 popc(nz & (a ^ s)). */
int qt_gemm_b1t2(const uint32_t* a_bits, int64_t lda_words, const uint32_t* w_nz, const uint32_t* w_sign,
                 int64_t ldw_words, int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, void* stream);

/* 8-bit codes x 8-bit codes -> int32.  backend: 0 = auto, 1 = tcgen05 (kind::i8, TMA + TMEM),
 * 2 = CUDA-core dp4a fallback (any shape).  lda/ldw in bytes; tcgen05 needs them % 16 == 0. */
int qt_gemm_i8(const void* a, int a_signed, int64_t lda, const void* w, int w_signed, int64_t ldw,
               int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, int backend, void* stream);

/* fp4 (e2m1) codes x fp4 (e2m1) codes -> exact integer accumulators, on tcgen05.mma kind::mxf4.block_scale with unit
 * scale factors (twice the MAC rate of kind::i8).  This is the 1-bit / ternary / 2-bit contraction of
 *   LinearBin / BinConv2d    binary_layers.py:42-46,103-106      (+-1 x +-1  ==  K - 2 popc(a ^ w))
 *   LinearTer / TerConv2d    terner_layers.py:47-51,89-92        ({-1,0,1} weights)
 *   LinearDorefa k <= 2      dorefa_layers.py:41-45
 * a: [M, lda] codes, w: [N, ldw] codes (two per byte; lda, ldw in ELEMENTS, multiples of 32; columns K.. zero).
 * Same integer epilogue as qt_gemm_i8.  tcgen05 only: QT_EUNSUPPORTED on anything but sm_100. */
int qt_gemm_f4(const void* a, int64_t lda, const void* w, int64_t ldw, int64_t M, int64_t N, int64_t K,
               const QtEpilogue* ep, void* stream);

/* Implicit-GEMM convolution (F.conv2d at binary_layers.py:105-106, terner_layers.py:91-92, dorefa_layers.py:79,81) on
 * channels-last 8-bit activation codes x_nhwc[B, H, W, C]: the A operand is fetched by TMA in im2col mode (no im2col
 * matrix is ever materialised; padding taps are zero-filled by the hardware), K runs over (kh, kw, c) with the channel
 * fastest, `w` is the expanded weight operand [N, kh*kw*C/groups] of conv group `group` (ldw bytes per row).
 * Same epilogue (use out_mode = 1 for NCHW output).  Returns QT_EUNSUPPORTED when (C/groups) % 32 != 0 or the device is
 * not sm_100: callers then use qt_im2col + qt_gemm_i8. */
typedef struct QtConvGeom {
  int64_t B, C, H, W;
  int kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int groups, group;
  int64_t OH, OW;
  /* optional explicit bounding box of the filter-window base positions along W (asymmetric padding), used when
     corner_mode != 0: positions lower_w .. (W - 1) + upper_w in steps of stride_w; pad_w is then ignored.  A W-folded conv
     (two adjacent pixels of a 64-byte-channel tensor viewed as one 128-byte pixel) runs as two such launches, one per
     output-column parity, with their own zero-padded filters (see engine.conv2d). */
  int corner_mode, lower_w, upper_w;
} QtConvGeom;

int qt_conv_i8(const void* x_nhwc, int a_signed, const QtConvGeom* g, const void* w, int w_signed, int64_t ldw,
               int64_t N, const QtEpilogue* ep, void* stream);

/* The same implicit GEMM on channels-last bf16 activations x_nhwc[B, H, W, C] and bf16 weights w[N, kh*kw*C/groups]
 * (tcgen05 kind::f16, fp32 accumulation): the first layer of every reference net, which sees the fp32 image
 * (models/Alexnet/Alexnet_Bin.py:13, models/Resnet/Resnet_bin.py:68, models/VGG/VGG_LinQuant.py:12 -> F.conv2d of a real
 * input with quantized weights).  Needs (C/groups) * 2 bytes % 32 == 0: callers feed the "plane pixel" tensor of
 * qt_image_planes, whose 16 slots per pixel hold the bf16 hi / mid / lo parts of the (<= 5) image channels, against weights
 * whose integer codes are repeated per part -- one pass, 24 significant bits of the input.  ldw in ELEMENTS. */
int qt_conv_bf16(const void* x_nhwc, const QtConvGeom* g, const void* w, int64_t ldw, int64_t N, const QtEpilogue* ep, void* stream);

/* fp32 NCHW image x[B, C, H, W] -> channels-last bf16 plane pixels with the conv's zero padding materialised:
 * pixel (h + pad_h, w + pad_w) of the padded [Hp, Wp] image holds 16 slots, slot p * C + c = part p of x[b, c, h, w]
 * (p = 0 hi, 1 mid, 2 lo; hi + mid + lo == x to 24 significant bits), every other slot / border pixel zero; planes * C <= 16.
 * Pixels of the source that fall outside [Hp, Wp] after the shift are dropped (a strided conv never reads them).
 * Space-to-depth: a fold_h x fold_w block of pixels is stored as ONE super pixel of fold_h * fold_w * 16 slots,
 *   out[b, hp / fold_h, wp / fold_w, ((hp % fold_h) * fold_w + wp % fold_w) * 16 + slot]     (fold_h = fold_w = 1: [B, Hp, Wp, 16]),
 * so that a stride-f filter of k taps per axis becomes a stride-1 filter of floor((k - 1) / f) + 1 super taps: a 7x7 / 2 stem
 * reads 4 x 4 taps of 128-byte super pixels instead of 49 taps of 32 bytes, an 11x11 / 4 stem (fold 1 x 4) 11 x 3 taps: whole
 * 128-byte k-blocks, a fraction of the TMA / MMA issues per output.  (Filters whose row fits a record use qt_image_windows,
 * which moves fewer bytes through L2.) */
int qt_image_planes(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int planes, int pad_h, int pad_w,
                    int64_t Hp, int64_t Wp, int fold_h, int fold_w, void* out, void* stream);

/* Row-window records of the same image, for filters whose whole row fits one record (kw * planes * C <= slots, slots * 2 bytes
 * = 32 / 64 / 128 / 256): out[b, hp, ow, slots] (bf16), slot kx * planes * C + p * C + c = part p of x[b, c, hp - pad_h,
 * ow * stride_w - pad_w + kx], zero outside the image and in the unused slots.  The conv over the record grid has kw = 1,
 * stride_w = 1 and one k-block per filter ROW: the 7x7 / 2 stem of models/Resnet/Resnet_bin.py:68 reads 7 x 128 bytes per
 * output pixel instead of the 16 x 128 bytes of its 2 x 2 space-to-depth form (63 of 64 slots used instead of 441 of 1024) --
 * the implicit GEMM is bound by L2 -> shared-memory bytes, not by the tensor pipe. */
int qt_image_windows(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int planes, int kw, int stride_w,
                     int pad_h, int pad_w, int64_t Hp, int64_t OW, int slots, void* out, void* stream);

/* fp32 classifier head of the residual nets (models/Resnet/Resnet_bin.py:104-107: avg_pool2d -> view -> nn.Linear):
 * out[b, n] = bias[n] + sum_c (mean over hw of x[b, hw, c]) * w[n, c], x channels-last [B, HW, C] (HW = 1: a plain fp32
 * Linear on [B, C]); bias may be NULL.  Every sum runs in a fixed order that depends on (HW, C) only, so a sample's logits do
 * not depend on the batch it is part of (library GEMMs choose split-K by batch size): the gathered logits of a sharded run
 * equal the single-GPU run bit for bit. */
int qt_head_f32(const float* x, int64_t B, int64_t HW, int64_t C, const float* w, int64_t ldw, const float* bias,
                int64_t N, float* out, int64_t ldo, void* stream);

/* Max-pool (nn.MaxPool2d, ceil_mode = False, dilation 1; OH = floor((H + 2 pad - k) / stride) + 1) on channels-last tensors.
 *   qt_pool_codes      8-bit activation codes [B, H, W, C] -> [B, OH, OW, C].  An activation quantizer is monotone, so
 *                      pool(quantize(clamp(bn(y)))) == quantize(clamp(bn(pool(y)))) whenever the BatchNorm scale of the channel
 *                      is >= 0, and equals the MIN-pool of the codes when it is negative: use_min[c] (optional, [C] bytes)
 *                      selects min for those channels.  This lets `conv -> MaxPool -> BatchNorm -> Hardtanh -> quantizer`
 *                      (models/Alexnet/Alexnet_Bin.py:13-22, benchmark/BinaryNet/AlexNetBin.py:13-24) run as conv with the
 *                      requant epilogue followed by a pool over 1-byte codes instead of fp32.  C % 16 == 0.
 *   qt_pool_quant_f32  fp32 [B, H, W, C] -> pooled fp32 (optional) and / or the 8-bit codes of the pooled values
 *                      (mode QT_Q_SIGN / QT_Q_TERNARY / QT_Q_DOREFA as qt_quant_act; codes_kind 1 int8, 2 uint8).  C % 4 == 0. */
typedef struct QtPoolGeom {
  int64_t B, H, W, C;
  int kh, kw, stride_h, stride_w, pad_h, pad_w;
  int64_t OH, OW;
} QtPoolGeom;

/* row_sum[r] = sum of the 8-bit codes of row r of codes[rows, ld] (zero padding included): the activation row sums the unsigned
 * DoReFa-8 weight zero point needs (see qt_gemm_i8) when a Linear layer reads the flattened codes a conv chain left. */
int qt_rowsum_codes(const void* codes, int is_unsigned, int64_t rows, int64_t ld, int32_t* row_sum, void* stream);

int qt_pool_codes(const void* x_nhwc, int is_unsigned, const QtPoolGeom* g, const uint8_t* use_min, void* out, void* stream);
int qt_pool_quant_f32(const float* x_nhwc, const QtPoolGeom* g, float* out, int mode, int bit_width, void* codes, int codes_kind,
                      int32_t* overflow, void* stream);

/* Per-output-pixel sum of the 8-bit codes under the filter window of conv group `group` (zero padding contributes 0):
 * row_sum[m] = sum_{kh,kw,c} x_nhwc[b, ih, iw, c].  It is the activation row sum the unsigned-weight (DoReFa-8) zero point
 * needs when the conv runs as an implicit GEMM and no im2col matrix exists.  `chan_sum` is int32 scratch of B*H*W elements. */
int qt_patch_rowsum(const void* x_nhwc, int is_unsigned, const QtConvGeom* g, int32_t* chan_sum, int32_t* row_sum, void* stream);

/* 16-bit float planes x 16-bit float planes -> fp32.  D = sum over passes p of A[pa[p]] . W[pw[p]]^T.
 * fmt 0: bf16, fmt 1: fp16 (both operands).  Planes are [M, lda] / [N, ldw] matrices `a_plane_stride` /
 * `w_plane_stride` elements apart.  backend as above (2 = CUDA-core fp32 FMA fallback). */
int qt_gemm_f16(const void* a, int64_t lda, int64_t a_plane_stride, const void* w, int64_t ldw,
                int64_t w_plane_stride, int fmt, int npass, const int* pa, const int* pw,
                int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, int backend, void* stream);

/* fp32 x fp32 CUDA-core GEMM (D = A . W^T), the always-available exact-fp32 route used for
 * ragged first layers; same epilogue. */
int qt_gemm_f32(const float* a, int64_t lda, const float* w, int64_t ldw,
                int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, void* stream);

/* ------------------------------------------------------------------------
 * Backward (straight-through estimators), SURVEY 8f-2.  The gradient contractions of the dense layers
 *   grad_x = g . W_q            (binary_connect.py:104-105, terner_connect.py:96-98, dorefa_connect.py:139-146, xnor_connect.py:121-122)
 *   grad_W = g^T . x  (+ STE)   (binary_connect.py:106-107, xnor_connect.py:123-126)
 * run on qt_gemm_f16 with bf16 hi/lo planes (relative error <= 2^-17); the operand contracted along its LEADING dimension
 * is transposed and split in one pass by qt_transpose_split:
 *   x fp32 [rows, ld_x] -> out bf16 [planes][cols, ld_out]  (out[p][c][r]; ld_out >= rows, multiple of 8 for the tensor path;
 *   columns rows..ld_out-1 zero-filled; plane stride = cols * ld_out).
 * qt_ste_clip: out = |x| <= thresh ? g : 0 -- the clip-mask STE of BinaryConnect / TernaryConnect
 * (binary_connect.py:30-38, terner_connect.py:29-34) in one pass.
 * ---------------------------------------------------------------------- */
int qt_transpose_split(const float* x, int64_t rows, int64_t cols, int64_t ld_x, void* out, int64_t ld_out, int planes,
                       void* stream);
int qt_ste_clip(const float* g, const float* x, float thresh, float* out, int64_t n, void* stream);

/* The one collective of the path (all-gather of the fp32 logits of a batch-sharded run) in push form: copy `bytes`
 * from `src` (local) to each of the `ndst` (<= 8) destinations -- pointers into PEER GPUs' gathered buffers, mapped into this
 * process (symmetric memory / cudaIpc) -- with 16-byte stores from `ctas` small CTAs (no shared memory, so they run beside the
 * persistent tcgen05 kernels of the next step).  Ends with a system-scope fence; the caller signals the peers afterwards
 * (stream-ordered barrier).  The reference has no collective at all (single process, `device.py:2`). */
int qt_peer_push(const void* src, void* const* dst, int ndst, int64_t bytes, int ctas, void* stream);

/* Tuning / test knobs (process-wide).  "f4_tile_n": qt_gemm_f4 tile width, 0 = auto, or 64 / 128 / 240.
 * Unknown names or values return QT_EINVAL. */
int qt_set_option(const char* name, int value);

/* Number of kernel launches issued by this library on the calling thread since the last reset
 * (bench.py's gpu_launches). */
int64_t qt_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* QTB200_H */
