/* gbcodec.h — C ABI of the B200-native heatmap codec (libgbcodec.so).
 *
 * The reference (MarkJhonBao/InfantPoseEstimation_GaussianBias) is pure Python
 * and has no FFI layer; its boundary for this path is a set of Python
 * callables.  Each entry point below replaces the arithmetic of one of them and
 * is what a binding added to the reference would call (see INTEGRATION.md for
 * the ctypes / torch.library stubs).  Citations are file:line in the reference.
 *
 * Conventions
 *   - every pointer named `d_*` or documented "device" is CUDA device memory on
 *     the current device; `h_*` is host memory.  All tensors are contiguous
 *     fp32 NCHW: heatmaps (B,K,H,W), offsets (B,K,2,H,W), coords (B,K,2).
 *   - `stream` is a cudaStream_t passed as void*.  Calls only enqueue work:
 *     they never allocate, never synchronise and keep no global state, so they
 *     are re-entrant and may be used from several host threads on different
 *     streams.  The caller owns every buffer, including the workspace.
 *   - return value: GBCODEC_OK (0) or a negative gbcodec_status.  Nothing throws.
 *     gbcodec_last_error() returns a thread-local description of the last failure.
 *   - W must be a multiple of 4 and every tensor 16-byte aligned (128-bit
 *     loads); 1 <= K <= GBCODEC_MAX_K; H*W <= GBCODEC_MAX_TILE (the six-term loss
 *     keeps two tiles in shared memory: H*W <= 28 000 there, GBCODEC_ERR_BAD_SHAPE above).
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef GBCODEC_H_
#define GBCODEC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GBCODEC_ABI_VERSION 3
#define GBCODEC_MAX_K 64            /* keypoint channels per image                  */
#define GBCODEC_MAX_PAIRS 64        /* limb pairs in the overlap term               */
#define GBCODEC_MAX_PARTNERS 4      /* limb pairs a single channel may take part in */
#define GBCODEC_MAX_TILE 32768      /* H*W                                           */

typedef enum gbcodec_status {
    GBCODEC_OK = 0,
    GBCODEC_ERR_NULL_POINTER = -1,
    GBCODEC_ERR_BAD_SHAPE = -2,      /* non-positive dims, K or tile too large, W % 4 != 0 */
    GBCODEC_ERR_UNALIGNED = -3,      /* a tensor pointer is not 16-byte aligned              */
    GBCODEC_ERR_BAD_ARGUMENT = -4,   /* bad mode / flags / skeleton / sigma                  */
    GBCODEC_ERR_WORKSPACE = -5,      /* workspace missing or too small                       */
    GBCODEC_ERR_CUDA = -6,           /* a CUDA runtime call failed (see gbcodec_last_error)  */
    GBCODEC_ERR_PEER_TIMEOUT = -7    /* an earlier sharded call gave up waiting for a peer: its results were NaN */
} gbcodec_status;

int gbcodec_abi_version(void);
const char* gbcodec_status_string(int status);
const char* gbcodec_last_error(void);

/* ---------------------------------------------------------------- encode ---
 * Gaussian target tiles + weights.  Replaces COCOPoseDataset._generate_target
 * (datasets/coco_dataset.py:185-250), batched: the reference runs it per sample
 * in DataLoader workers and ships the tiles over PCIe (train.py:160); here the
 * tiles are written straight into HBM from (B,K,2) keypoints and (B,K) flags.
 *   d_kps    (B,K,2) input-image pixels       d_vis    (B,K) visibility {0,1,2}
 *   d_target (B,K,H,W) out                     d_weight (B,K) out (= (B,K,1))
 *   in_w,in_h  network input size; sigma  Gaussian sigma in heatmap pixels.
 */
int gbcodec_encode_f32(const float* d_kps, const float* d_vis, float* d_target, float* d_weight,
                       int B, int K, int H, int W, float in_w, float in_h, double sigma, void* stream);

/* ---------------------------------------------------------------- decode ---
 * Replaces HeatmapRegressionHead.decode (models/fusion_head.py:309-365) —
 * global soft-argmax (:37-71), local (2r+1)^2 softmax centroid around the
 * rounded soft-argmax (:84-128), sigmoid(alpha) blend (:151-172), bilinear
 * offset-map correction (:342-363) — and, when d_hm_flipped is given, the
 * flip-test average of PoseEstimator.inference (models/pose_estimator.py:
 * 303-319) fused in front: the tile decoded is
 *   (hm[b,k,y,x] + hm_flipped[b,perm[k],y,W-1-x]) / 2.
 *   d_hm           (B,K,H,W)
 *   d_hm_flipped   (B,K,H,W) raw head output for the W-flipped image, or NULL
 *   d_flip_perm    K int32, channel permutation of the flip pairs (NULL = identity)
 *   d_off          (B,K,2,H,W), required with GBCODEC_DECODE_APPLY_OFFSET; may also be a device-accessible
 *                  (pinned, mapped) HOST pointer: 8 taps per tile are read, in place
 *   d_alpha_param  device scalar, raw learnable alpha (sigmoid applied here), required with REFINE
 *   d_fusion_weight device scalar; already sigmoid-ed as the head publishes it
 *                  (fusion_head.py:306) unless GBCODEC_DECODE_FUSION_WEIGHT_RAW
 *   d_coords (B,K,2) out, heatmap pixels     d_scores (B,K) out, raw tile maximum
 *   d_centre (B,K,2) int32 out or NULL: the rounded window centre (integer peak index)
 */
#define GBCODEC_DECODE_REFINE            1u
#define GBCODEC_DECODE_APPLY_OFFSET      2u
#define GBCODEC_DECODE_FUSION_WEIGHT_RAW 4u
int gbcodec_decode_f32(const float* d_hm, const float* d_hm_flipped, const int32_t* d_flip_perm,
                       const float* d_off, const float* d_alpha_param, const float* d_fusion_weight,
                       int B, int K, int H, int W, int local_radius, unsigned flags,
                       float* d_coords, float* d_scores, int32_t* d_centre, void* stream);

/* Replaces PoseEstimator.decode_heatmaps (models/pose_estimator.py:331-373) and
 * the arg-max family of utils/postprocess.py:10-184.  First maximum wins ties.
 *   mode GBCODEC_ARGMAX_PLAIN    integer peak                         (postprocess.py:10-34, shift=False)
 *        GBCODEC_ARGMAX_QUARTER  +0.25*sign(neighbour difference)     (pose_estimator.py:361-371)
 *        GBCODEC_ARGMAX_TAYLOR   second-order sub-pixel step           (postprocess.py:37-75)
 *   d_coords (B,K,2) out   d_maxvals (B,K) out   d_index (B,K) int32 out or NULL (flat y*W+x)
 */
#define GBCODEC_ARGMAX_PLAIN   0
#define GBCODEC_ARGMAX_QUARTER 1
#define GBCODEC_ARGMAX_TAYLOR  2
int gbcodec_decode_argmax_f32(const float* d_hm, int B, int K, int H, int W, int mode,
                              float* d_coords, float* d_maxvals, int32_t* d_index, void* stream);

/* Replaces utils/postprocess.py:138-184 (coordinate_refinement): linear-weight
 * centroid of the window around trunc(coords).  d_coords_in/out (B,K,2). */
int gbcodec_refine_centroid_f32(const float* d_hm, const float* d_coords_in, int B, int K, int H, int W,
                                int window, float* d_coords_out, void* stream);

/* ------------------------------------------------------------------ loss ---
 * Replaces FusionPoseLoss.forward + its autograd backward
 * (models/fusion_head.py:745-806, terms :637-743 and :405-559).
 */
typedef struct gbcodec_loss_desc {
    int32_t B, K, H, W;
    float   in_w, in_h;          /* network input size (fusion_head.py:679-680)                     */
    float   lambdas[6];          /* heatmap, offset, peak, variance, overlap, shape (:608-618)      */
    double  target_sigma;        /* sigma of the variance / entropy targets (:467, :550)           */
    double  encode_sigma;        /* sigma of the on-the-fly target tiles (d_target == NULL)        */
    int32_t use_target_weight;   /* :651, :706, :737 — terms 1-3 only                                */
    int32_t n_pairs;             /* limb pairs (i,j), i,j < K (:389-394, :503-505)                  */
    int32_t pairs[GBCODEC_MAX_PAIRS][2];
} gbcodec_loss_desc;

size_t gbcodec_loss_workspace_bytes(int B, int K, int H, int W);

/* Batch normalisers WITHOUT their epsilons: out[0] = sum(w), out[1] = sum over limbs
 * of w_i*w_j (fusion_head.py:480,523-527).  With target_given == 0 the weights first
 * go through the encoder's off-map rule (coco_dataset.py:227-229), which is what the
 * on-the-fly mode of the loss uses.  A batch-sharded job all-reduces the two floats
 * over its ranks and hands the result to the loss as d_denoms. */
int gbcodec_loss_denominators_f32(const gbcodec_loss_desc* desc, const float* d_weight, const float* d_gt_kps,
                                  int target_given, float* d_out2_raw_sums,
                                  void* d_workspace, size_t workspace_bytes, void* stream);

/* One pass over the batch: the seven lambda-weighted loss scalars and, when the
 * grad pointers are given, d(total_loss)/d(hm, off, var) scaled by *d_grad_scale —
 * every heatmap tile is read from HBM once.
 *   d_hm (B,K,H,W)  d_off (B,K,2,H,W)  d_var (B,K,H,W) or NULL
 *            d_off may also be a device-accessible (pinned, mapped) HOST pointer: the pass reads
 *            at most 16 floats per tile from it, so a caller whose offset maps live in host
 *            memory need not ship them over PCIe
 *   d_target (B,K,H,W), or NULL: tiles are generated on the fly from d_gt_kps and
 *            d_weight exactly as gbcodec_encode_f32 would (no HBM traffic for them)
 *   d_weight (B,K)   d_gt_kps (B,K,2) input-image pixels
 *   d_denoms  NULL, or 2 device floats: raw sums (no epsilon) of w and of w_i*w_j over the
 *            GLOBAL batch — a rank holding a shard passes the all-reduced values
 *   d_grad_scale NULL (=1) or device scalar: the upstream gradient of total_loss
 *   d_losses7 out: heatmap, offset, peak, variance, overlap, shape, total
 *   d_grad_hm/off/var out (all NULL = forward only; d_grad_var NULL iff d_var NULL)
 */
int gbcodec_fusion_loss_f32(const gbcodec_loss_desc* desc,
                            const float* d_hm, const float* d_off, const float* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps,
                            const float* d_denoms, const float* d_grad_scale,
                            float* d_losses7, float* d_grad_hm, float* d_grad_off, float* d_grad_var,
                            void* d_workspace, size_t workspace_bytes, void* stream);

/* Same pass with the keypoint decode of gbcodec_decode_f32 (no flip) folded in:
 * the "fused step" encode + loss fwd/bwd + decode of BASELINE.json. */
int gbcodec_fusion_step_f32(const gbcodec_loss_desc* desc,
                            const float* d_hm, const float* d_off, const float* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps,
                            const float* d_denoms, const float* d_grad_scale,
                            float* d_losses7, float* d_grad_hm, float* d_grad_off, float* d_grad_var,
                            const float* d_alpha_param, const float* d_fusion_weight, int local_radius, unsigned decode_flags,
                            float* d_coords, float* d_scores,
                            void* d_workspace, size_t workspace_bytes, void* stream);

/* gbcodec_fusion_step_f32 for a head that reduces its variance branch to the per-tile mean itself (the Softplus +
 * mean_N fused into the epilogue of its last convolution, fusion_head.py:245-251): the loss uses V only through
 * mean_N(V) (:467-478), so the map need not exist.  d_var_mean (B,K) replaces d_var, d_grad_var_mean (B,K) =
 * d(total)/d(mean_N(V)) replaces d_grad_var; d(total)/dV_i = d_grad_var_mean / N if the caller needs it.
 * Algorithmic HBM bytes per tile: 16N instead of 24N.  Tile shapes 64x48, 64x64, 96x72, 128x128.  d_coords/d_scores may both
 * be NULL (loss only); gradients all given or all NULL. */
int gbcodec_fusion_step_vmean_f32(const gbcodec_loss_desc* desc,
                            const float* d_hm, const float* d_off, const float* d_var_mean, const float* d_target,
                            const float* d_weight, const float* d_gt_kps,
                            const float* d_denoms, const float* d_grad_scale,
                            float* d_losses7, float* d_grad_hm, float* d_grad_off, float* d_grad_var_mean,
                            const float* d_alpha_param, const float* d_fusion_weight, int local_radius, unsigned decode_flags,
                            float* d_coords, float* d_scores,
                            void* d_workspace, size_t workspace_bytes, void* stream);

/* The caller's half of the per-tile-mean path: the variance branch's last convolution hands over its RAW output (the
 * Softplus module of models/fusion_head.py:245-251 dropped) and gets mean_N(softplus(raw)) per tile — what
 * gbcodec_fusion_step_vmean_f32 takes as d_var_mean; the step's d_grad_var_mean goes back through the second call as the
 * gradient of the raw map, d raw_i = g_tile / N * sigmoid(raw_i) (torch.nn.Softplus, beta 1, threshold 20).  The
 * (B,K,H,W) variance map and its gradient map never exist.  d_raw (B,K,H,W) 16-byte aligned, d_mean / d_grad_mean (B,K). */
int gbcodec_softplus_mean_f32(const float* d_raw, float* d_mean, int B, int K, int H, int W, void* stream);
int gbcodec_softplus_mean_backward_f32(const float* d_raw, const float* d_grad_mean, float* d_grad_raw,
                                       int B, int K, int H, int W, void* stream);

/* Backward for an arbitrary upstream gradient on the seven outputs (the autograd backward of train.py:182 for
 * whatever reaches the loss dict of fusion_head.py:795-806).  The gradients written by the forward assume
 * d(total)=*d_grad_scale and nothing on the six terms.  This call reads the actual upstream vector on the device and
 *   - returns immediately inside the kernels if the stored gradients already carry it,
 *   - rescales the three gradient tensors in place if they are off by one common factor,
 *   - otherwise recomputes them with per-term weights.
 * No host synchronisation in any case (three launches, the usual answer costs a few microseconds).
 *   d_held6   NULL, or 6 device floats the caller keeps with the stored gradients: the upstream factor each term's
 *             share of them carries.  The call compares against it (held_valid != 0) or against *d_grad_scale
 *             (held_valid == 0: first backward after the forward) and writes the new factors back, so that a second
 *             backward through the same stored gradients (retain_graph; per-term losses, then total_loss) is right.
 *             With NULL every call assumes the forward's state: correct for one backward per forward only.
 *   workspace_from_forward != 0: d_workspace is the forward's workspace, untouched since; the weights, patch geometry
 *             and normalisers in it are reused instead of being computed again. */
int gbcodec_fusion_loss_backward_f32(const gbcodec_loss_desc* desc,
                            const float* d_hm, const float* d_off, const float* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps,
                            const float* d_denoms, const float* d_grad_scale, const float* d_grad_losses7,
                            float* d_grad_hm, float* d_grad_off, float* d_grad_var,
                            float* d_held6, int held_valid, int workspace_from_forward,
                            void* d_workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------- batch-sharded jobs ---
 * One process per GPU, each rank holding whole images.  The data path needs no exchange; the
 * loss needs two scalar ones per step: the batch normalisers before the tile
 * kernel and the seven loss scalars after it.  Two ways to do them:
 *   (a) portable: gbcodec_loss_denominators_f32 -> all-reduce (NCCL) -> d_denoms of
 *       gbcodec_fusion_step_f32 -> all-reduce of d_losses7;
 *   (b) gbcodec_fusion_step_sharded_f32: the kernels that produce those scalars write them into
 *       every peer's mailbox over NVLink / NVSwitch (cudaIpc-mapped device memory) and read the
 *       peers' values from their own — no NCCL call, no extra launch, fixed rank order (every
 *       rank gets the same bits).  d_losses7 then holds the GLOBAL losses on every rank and the
 *       gradients are this rank's share of the global-batch gradients.
 * Set-up (these three calls allocate / map / free device memory and are NOT stream-ordered):
 *   gbcodec_peer_create   allocates this rank's mailbox, returns its IPC handle (64 bytes)
 *   gbcodec_peer_connect  maps the peers' mailboxes; all_handles = world x 64 bytes in rank order
 *                         (gathered by the host program, e.g. torch.distributed.all_gather)
 *   gbcodec_peer_destroy
 * Every rank must make the same sequence of sharded calls.  A rank that waits for a peer longer than the
 * time-out (gbcodec_peer_set_timeout, default 120 s — a peer that saves a checkpoint or evaluates is late, not
 * dead) gives up: the kernel writes NaN into the normalisers / losses it could not complete (nothing is ever
 * computed from a stale mailbox slot), raises a sticky flag in mapped host memory, and the NEXT sharded call of
 * that rank returns GBCODEC_ERR_PEER_TIMEOUT without launching anything.  gbcodec_peer_status reads the count
 * (synchronises).  A call that fails its argument checks does not advance the exchange's sequence number.
 */
#define GBCODEC_PEER_HANDLE_BYTES 64
#define GBCODEC_MAX_PEERS 16
int gbcodec_peer_create(int rank, int world, void** ctx_out, unsigned char* handle_out);
int gbcodec_peer_connect(void* ctx, const unsigned char* all_handles);
int gbcodec_peer_status(void* ctx, int* h_timeouts);
int gbcodec_peer_set_timeout(void* ctx, double seconds);
int gbcodec_peer_destroy(void* ctx);

/* gbcodec_fusion_step_f32 for a rank of a sharded job.  d_coords/d_scores may both be NULL (loss
 * only).  d_denoms_out: NULL or 2 device floats that receive the global raw sums (what the
 * backward of gbcodec_fusion_loss_backward_f32 takes as d_denoms).
 *
 * Both exchanges can be taken off the step's critical path (the normalisers depend on the visibility flags and the
 * keypoints only — fusion_head.py:480,523-527 — and the seven losses are logged, nothing is computed from them):
 *   d_denoms_global  NULL: the normalisers are exchanged inside this call (the tile kernel waits for every peer).
 *                    Else 2 device floats, the GLOBAL raw sums, e.g. from gbcodec_peer_denominators_f32 issued on a side
 *                    stream while the previous step ran: no exchange in front of the tile kernel.
 *   defer_losses     != 0 (needs d_denoms_global): the step writes its six terms into every peer's mailbox and does not
 *                    wait for theirs; d_losses7 receives THIS rank's share of the global losses, and
 *                    gbcodec_peer_collect_losses_f32(ctx, steps_back, ...) adds the ranks' shares of the step made
 *                    `steps_back` (0, 1 or 2) sharded calls ago, in rank order (same bits on every rank).  Collect within
 *                    two steps: the mailbox keeps four steps of terms. */
int gbcodec_fusion_step_sharded_f32(const gbcodec_loss_desc* desc,
                            const float* d_hm, const float* d_off, const float* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps, const float* d_grad_scale,
                            float* d_losses7, float* d_grad_hm, float* d_grad_off, float* d_grad_var,
                            const float* d_alpha_param, const float* d_fusion_weight, int local_radius, unsigned decode_flags,
                            float* d_coords, float* d_scores, float* d_denoms_out,
                            const float* d_denoms_global, int defer_losses,
                            void* d_workspace, size_t workspace_bytes, void* peer_ctx, void* stream);
/* gbcodec_loss_denominators_f32 + the exchange: d_out2_global_sums receives the sums over ALL ranks.  Every rank must
 * call it once per step, in step order (it has its own sequence numbers and may run ahead of the steps by one). */
int gbcodec_peer_denominators_f32(const gbcodec_loss_desc* desc, const float* d_weight, const float* d_gt_kps, int target_given,
                                  float* d_out2_global_sums, void* d_workspace, size_t workspace_bytes, void* peer_ctx, void* stream);
int gbcodec_peer_collect_losses_f32(void* peer_ctx, int steps_back, float* d_losses7, void* stream);

/* ------------------------------------------------- Gen-B family (next rows) ---
 * The reference carries a second generation of the same codec ("Gen-B":
 * data/, models/losses.py, utils/postprocess.py).  Same tile layout, same rules.
 */

/* Alternate encoders.
 *   GBCODEC_ENCODE_PATCH          COCOPoseDataset._generate_target (datasets/coco_dataset.py:185-250);
 *                                 what gbcodec_encode_f32 does
 *   GBCODEC_ENCODE_PATCH_CLIPPED  PreemieCocoDataset._generate_heatmaps (data/coco_dataset.py:222-287):
 *                                 mu in float32, weight binarised to 1, the joint must lie inside the
 *                                 map (:250), and the patch origin is clamped to 0 BEFORE the patch
 *                                 slice is derived (:262-263, :277) — for mu < 3 sigma the patch's own
 *                                 top-left corner lands on pixel 0
 *   GBCODEC_ENCODE_DENSE          GenerateTarget (data/pose_transforms.py:385-457): sub-pixel centre,
 *                                 exp over the whole tile, weight 1/0 (visible and inside the map)
 * d_vis (B,K): > 0 means visible for the two Gen-B modes. */
#define GBCODEC_ENCODE_PATCH         0
#define GBCODEC_ENCODE_PATCH_CLIPPED 1
#define GBCODEC_ENCODE_DENSE         2
int gbcodec_encode_mode_f32(const float* d_kps, const float* d_vis, float* d_target, float* d_weight,
                            int B, int K, int H, int W, float in_w, float in_h, double sigma, int mode, void* stream);

/* postprocess_predictions (utils/postprocess.py:296-340) in ONE pass over the heatmaps:
 * fused_decode (:78-135: Taylor arg-max, optional scale to the 256-px image, confidence-adaptive
 * blend with the regression branch) -> coordinate_refinement (:138-184) -> filter_low_confidence
 * (:226-238) -> transform_preds (:270-292).  Every stage is optional, so the same entry point
 * also serves fused_decode alone.
 *   d_regression (B,K,2) or NULL.  The reference rescales it by image_size when its batch-wide
 *                maximum is <= 1.0 (:119); that test runs on the device (d_workspace: 16 bytes).
 *   d_center, d_scale (B,2) or NULL; required with scale_to_image or transform.
 *   d_preds (B,K,2) out   d_maxvals (B,K) out   d_mask (B,K) out or NULL
 */
typedef struct gbcodec_postprocess_desc {
    int32_t B, K, H, W;
    int32_t argmax_mode;      /* GBCODEC_ARGMAX_*; fused_decode uses TAYLOR                          */
    int32_t scale_to_image;   /* preds *= image_size / (W, H)            (postprocess.py:105-114)    */
    float   image_size;       /* 256 in the reference                    (:108, :121)                */
    int32_t refine_window;    /* coordinate_refinement window, 0 = skip  (:138)                      */
    int32_t filter;           /* filter_low_confidence on/off            (:226)                      */
    float   threshold;
    int32_t transform;        /* transform_preds on/off                  (:270)                      */
    float   input_w, input_h; /* its input_size (default 256, 256)                                    */
} gbcodec_postprocess_desc;
int gbcodec_postprocess_f32(const gbcodec_postprocess_desc* desc, const float* d_hm, const float* d_regression,
                            const float* d_center, const float* d_scale,
                            float* d_preds, float* d_maxvals, float* d_mask, void* d_workspace16, void* stream);

/* Heatmap pixels -> input pixels -> original image (validate.py:31-36,102-119; inference.py:143-175):
 *   c_in = c * (in / hm);   c_img = c_in / in * scale + center - scale / 2
 * in exactly that order of float32 operations.  d_coords_in/out (B,K,2) (may alias), d_center/d_scale (B,2). */
int gbcodec_coords_to_image_f32(const float* d_coords_in, const float* d_center, const float* d_scale,
                                int B, int K, int H, int W, float in_w, float in_h, float* d_coords_out, void* stream);

/* CombinedLoss (models/losses.py:205-290) and its parts, forward + backward in one pass:
 *   heatmap    FusedPoseLoss (:10-47): mean_{B,K,H,W}(crit(p,t) * w)
 *              with GBCODEC_CRIT_MSE_WEIGHTED: mean((p*w - t*w)^2) — KeypointMSELoss
 *              (models/pose_estimator.py:102-143) and, with heatmap_scale = 0.5, JointsMSELoss (:174-202)
 *   morph      MorphologyShapeLoss (:50-135): spatial mean / variance of pred and target
 *   regression, refined   OffsetRegressionLoss (:138-171) on (B,K,2) coordinates
 *   total = w_heatmap*heatmap + w_morph*morph + w_reg*(regression + refined)
 * Algorithmic HBM bytes per tile: read pred, target (8N); write d_pred (4N).
 */
#define GBCODEC_CRIT_MSE          0
#define GBCODEC_CRIT_SMOOTHL1     1
#define GBCODEC_CRIT_L1           2   /* coordinates only */
#define GBCODEC_CRIT_MSE_WEIGHTED 3   /* heatmaps only    */
#define GBCODEC_TERM_HEATMAP    1u
#define GBCODEC_TERM_MORPH      2u
#define GBCODEC_TERM_REGRESSION 4u
#define GBCODEC_TERM_REFINED    8u
typedef struct gbcodec_combined_desc {
    int32_t  B, K, H, W;
    int32_t  norm_batch;          /* batch size in the means' denominators; 0 = B.  A rank holding a
                                     shard passes the global batch and all-reduces the five scalars  */
    uint32_t terms;               /* GBCODEC_TERM_* present in `predictions` / `targets`             */
    int32_t  heatmap_criterion;   /* GBCODEC_CRIT_MSE | _SMOOTHL1 | _MSE_WEIGHTED                     */
    int32_t  coord_criterion;     /* GBCODEC_CRIT_SMOOTHL1 | _L1 | _MSE                               */
    int32_t  use_target_weight;   /* heatmap term only (:41); the other terms use w whenever given   */
    float    heatmap_scale;       /* 1, or 0.5 for JointsMSELoss (:196)                               */
    float    lambda_variance, lambda_mean;   /* MorphologyShapeLoss (:66-69)                          */
    float    w_heatmap, w_morph, w_reg;      /* CombinedLoss (:228-231)                               */
} gbcodec_combined_desc;

size_t gbcodec_combined_workspace_bytes(int B, int K);

/*   d_pred, d_target (B,K,H,W) — required with TERM_HEATMAP / TERM_MORPH
 *   d_weight (B,K) or NULL
 *   d_coords, d_refined, d_target_coords (B,K,2) — with TERM_REGRESSION / TERM_REFINED
 *   d_grad_scale NULL (=1) or device scalar: upstream gradient of `total` assumed by the forward
 *   d_losses5 out: heatmap, morph, regression, refined, total (absent terms are 0)
 *   d_grad_pred (B,K,H,W), d_grad_coords, d_grad_refined (B,K,2): out, each NULL = not wanted
 */
int gbcodec_combined_loss_f32(const gbcodec_combined_desc* desc,
                              const float* d_pred, const float* d_target, const float* d_weight,
                              const float* d_coords, const float* d_refined, const float* d_target_coords,
                              const float* d_grad_scale, float* d_losses5,
                              float* d_grad_pred, float* d_grad_coords, float* d_grad_refined,
                              void* d_workspace, size_t workspace_bytes, void* stream);

/* Backward for an arbitrary upstream gradient on the five outputs: returns inside the kernels if it
 * equals what the forward assumed (d(total) = *d_grad_scale, nothing on the four terms), otherwise
 * recomputes the gradients with per-term weights.  No host synchronisation. */
int gbcodec_combined_loss_backward_f32(const gbcodec_combined_desc* desc,
                              const float* d_pred, const float* d_target, const float* d_weight,
                              const float* d_coords, const float* d_refined, const float* d_target_coords,
                              const float* d_grad_scale, const float* d_grad_losses5,
                              float* d_grad_pred, float* d_grad_coords, float* d_grad_refined,
                              void* d_workspace, size_t workspace_bytes, void* stream);

/* The same two calls for float16 predictions (training under autocast: the head's heatmaps are half, the data
 * loader's targets float32 — torch's autocast runs mse_loss / smooth_l1_loss on inputs cast to float32 and casts the
 * gradient back, models/losses.py:10-47 under train.py's autocast).  d_pred_f16 and d_grad_pred_f16 are (B,K,H,W)
 * float16, every other tensor is float32 as above; the values are up-cast where they are read and the gradient is
 * rounded to half once, AFTER it has met the upstream factor (*d_grad_scale in the forward, the actual upstream in
 * the backward, which recomputes only if the two differ).  The call must contain the heatmap term.
 * Algorithmic HBM bytes per tile: read pred (2N) + target (4N), write d_pred (2N). */
int gbcodec_combined_loss_f16(const gbcodec_combined_desc* desc,
                              const void* d_pred_f16, const float* d_target, const float* d_weight,
                              const float* d_coords, const float* d_refined, const float* d_target_coords,
                              const float* d_grad_scale, float* d_losses5,
                              void* d_grad_pred_f16, float* d_grad_coords, float* d_grad_refined,
                              void* d_workspace, size_t workspace_bytes, void* stream);
int gbcodec_combined_loss_backward_f16(const gbcodec_combined_desc* desc,
                              const void* d_pred_f16, const float* d_target, const float* d_weight,
                              const float* d_coords, const float* d_refined, const float* d_target_coords,
                              const float* d_grad_scale, const float* d_grad_losses5,
                              void* d_grad_pred_f16, float* d_grad_coords, float* d_grad_refined,
                              void* d_workspace, size_t workspace_bytes, void* stream);

/* The plain heatmap head (PoseEstimator with head_type='heatmap') in ONE pass over the heatmaps:
 * KeypointMSELoss forward + backward (models/pose_estimator.py:102-143: mean((p w - t w)^2) over B*K*H*W),
 * the target tiles generated on the fly as COCOPoseDataset._generate_target builds them
 * (datasets/coco_dataset.py:185-250) unless d_target is given, and decode_heatmaps (:331-373) of the
 * same tiles.  Algorithmic HBM bytes per tile: read hm (4N) + write d_hm (4N).
 *   d_target NULL: tiles and weights come from d_gt_kps (B,K,2) and d_weight = visibility (B,K);
 *            given: d_weight (B,K) or NULL is used as is
 *   norm_batch  batch size in the mean's denominator; 0 = B (a rank holding a shard passes the global batch)
 *   d_loss out, 1 float   d_grad_hm (B,K,H,W) out or NULL   d_grad_scale NULL (=1) or device scalar
 *   d_coords (B,K,2), d_maxvals (B,K), d_index (B,K) int32: out or NULL — gbcodec_decode_argmax_f32's outputs
 *   workspace: gbcodec_combined_workspace_bytes(B, K)
 */
int gbcodec_heatmap_step_f32(const float* d_hm, const float* d_target, const float* d_weight, const float* d_gt_kps,
                             int B, int K, int H, int W, float in_w, float in_h, double sigma,
                             int use_target_weight, int norm_batch, const float* d_grad_scale,
                             float* d_loss, float* d_grad_hm, int argmax_mode,
                             float* d_coords, float* d_maxvals, int32_t* d_index,
                             void* d_workspace, size_t workspace_bytes, void* stream);

/* float16 maps — what the head hands the loss under autocast (train.py:171).
 * d_hm, d_off, d_var and the three gradients are IEEE binary16 (same shapes); losses, coordinates, scores, weights,
 * keypoints and target tiles stay float32.  Values are up-cast where they enter the kernel and every sum runs in
 * float32, so the losses and the decode are exactly those of the float32 entry points on the up-cast maps, at half the
 * bytes.  A gradient has to meet its upstream factor — the loss scaler's 2^16 — BEFORE it is rounded to half, or it
 * underflows.  So:
 *   gbcodec_fusion_step_f16           stores gradients only if the gradient pointers are given, pre-multiplied by
 *                                     *d_grad_scale: the upstream factor the caller EXPECTS (last step's loss scale);
 *   gbcodec_fusion_loss_backward_f16  compares the actual upstream d_grad_losses7 with that assumption on the device
 *                                     and returns inside the kernel if it held (gradients_stored != 0); otherwise, or if
 *                                     nothing was stored, it computes the gradients with the actual per-term factors.
 * In steady state a training step is ONE pass (12N bytes per tile instead of 24N), a changed loss scale costs one more.
 * Tile shapes 64x48, 64x64, 96x72, 128x128 (GBCODEC_ERR_BAD_SHAPE otherwise: up-cast and use the float32 entry points).
 * Pixels whose logit equals a limb partner's exactly (frequent in half precision) take their half of the overlap gradient
 * in a second store: those gradients are rounded twice (<= 1 ulp of half). */
int gbcodec_fusion_step_f16(const gbcodec_loss_desc* desc,
                            const void* d_hm, const void* d_off, const void* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps, const float* d_denoms, const float* d_grad_scale,
                            float* d_losses7, void* d_grad_hm, void* d_grad_off, void* d_grad_var,
                            const float* d_alpha_param, const float* d_fusion_weight, int local_radius, unsigned decode_flags,
                            float* d_coords, float* d_scores, void* d_workspace, size_t workspace_bytes, void* stream);
int gbcodec_fusion_loss_backward_f16(const gbcodec_loss_desc* desc,
                            const void* d_hm, const void* d_off, const void* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps, const float* d_denoms,
                            const float* d_grad_scale, int gradients_stored, const float* d_grad_losses7,
                            void* d_grad_hm, void* d_grad_off, void* d_grad_var,
                            float* d_held6, int held_valid, int workspace_from_forward,
                            void* d_workspace, size_t workspace_bytes, void* stream);

/* Kernels launched by this library in this process so far (every launch site counts itself; memsets and copies are not
 * kernels).  bench.py reads it around its timed region for the `gpu_launches` it reports. */
unsigned long long gbcodec_launch_count(void);

/* Measurement hook (bench.py): the next gbcodec_fusion_loss_f32 / _step_f32 calls made
 * by THIS host thread record `start_event` right before and `stop_event` right after
 * the per-tile loss kernel, on the stream of the call.  Both are cudaEvent_t passed as
 * void*; pass NULL, NULL to switch the hook off.  Thread-local, no other state. */
int gbcodec_profile_loss_kernel(void* start_event, void* stop_event);

#ifdef __cplusplus
}
#endif
#endif /* GBCODEC_H_ */
