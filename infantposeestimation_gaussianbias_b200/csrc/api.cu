// api.cu — the C ABI declared in include/gbcodec.h: argument checks, status
// mapping and launches.  No torch types, no allocation, no synchronisation.
#include <atomic>
#include <stdlib.h>
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace gbc {

static thread_local char g_last_error[512] = "";

int fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return status;
}

// GBCODEC_SYNC_DEBUG=1: wait for the device after every launch and report the kernel that faulted (debugging only:
// it serialises the host with the device and defeats programmatic dependent launch)
static bool sync_debug() {
    static const int v = [] { const char* e = getenv("GBCODEC_SYNC_DEBUG"); return (e && e[0] == '1') ? 1 : 0; }();
    return v != 0;
}
static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && sync_debug()) {
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) fprintf(stderr, "[gbcodec] %s faulted: %s\n", what, cudaGetErrorString(e));
    }
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return GBCODEC_OK;
}

// encode.cu / decode.cu / loss.cu
int launch_encode(const float*, const float*, float*, float*, int, int, int, int, float, float, double, int, cudaStream_t);
int launch_decode(const float*, const float*, const int32_t*, const float*, const float*, const float*, int, int, int, int,
                  int, unsigned, float*, float*, int32_t*, cudaStream_t);
int launch_argmax(const float*, int, int, int, int, int, float*, float*, int32_t*, cudaStream_t);
int launch_centroid(const float*, const float*, int, int, int, int, int, float*, cudaStream_t);
size_t loss_workspace_bytes(int B, int K);
int loss_denominators(const gbcodec_loss_desc*, const float*, const float*, int, float*, void*, size_t, cudaStream_t);
int fusion_loss(const gbcodec_loss_desc*, const float*, const float*, const float*, const float*, const float*, const float*,
                const float*, const float*, float*, float*, float*, float*,
                const float*, const float*, int, unsigned, float*, float*, void*, size_t, cudaStream_t, void*, float*, int, const float*, float*, int = 0);
int launch_softplus_mean(const float*, float*, int, int, int, int, cudaStream_t);
int launch_softplus_mean_backward(const float*, const float*, float*, int, int, int, int, cudaStream_t);
int peer_denominators(const gbcodec_loss_desc*, const float*, const float*, int, float*, void*, size_t, void*, cudaStream_t);
int peer_collect_losses(void*, int, float*, cudaStream_t);
int peer_create(int, int, void**, unsigned char*);
int peer_connect(void*, const unsigned char*);
int peer_status(void*, int*);
int peer_destroy(void*);
int peer_set_timeout(void*, double);
int fusion_loss_backward(const gbcodec_loss_desc*, const float*, const float*, const float*, const float*, const float*,
                         const float*, const float*, const float*, const float*, float*, float*, float*, void*, size_t,
                         cudaStream_t, int, int, float*, int, int);

int launch_postprocess(const gbcodec_postprocess_desc*, const float*, const float*, const float*, const float*, float*, float*,
                       float*, void*, cudaStream_t);
int launch_coords_to_image(const float*, const float*, const float*, int, int, int, int, float, float, float*, cudaStream_t);
size_t combined_workspace_bytes(int B, int K);
int combined_loss(const gbcodec_combined_desc*, const float*, const float*, const float*, const float*, const float*, const float*,
                  const float*, float*, float*, float*, float*, void*, size_t, cudaStream_t, int half_io = 0);
int combined_loss_backward(const gbcodec_combined_desc*, const float*, const float*, const float*, const float*, const float*,
                           const float*, const float*, const float*, float*, float*, float*, void*, size_t, cudaStream_t, int half_io = 0);

int heatmap_step(const float*, const float*, const float*, const float*, int, int, int, int, float, float, double, int, int,
                 const float*, float*, float*, int, float*, float*, int32_t*, void*, size_t, cudaStream_t);

void set_profile_events(cudaEvent_t, cudaEvent_t);

static int check_tile_shape(const char* who, int B, int K, int H, int W) {
    if (B <= 0 || K <= 0 || H <= 0 || W <= 0) return fail(GBCODEC_ERR_BAD_SHAPE, "%s: B,K,H,W must be positive (got %d,%d,%d,%d)", who, B, K, H, W);
    if (K > GBCODEC_MAX_K) return fail(GBCODEC_ERR_BAD_SHAPE, "%s: K=%d exceeds %d", who, K, GBCODEC_MAX_K);
    if (W % 4) return fail(GBCODEC_ERR_BAD_SHAPE, "%s: W=%d is not a multiple of 4", who, W);
    if ((long long)H * W > GBCODEC_MAX_TILE) return fail(GBCODEC_ERR_BAD_SHAPE, "%s: tile %dx%d exceeds %d pixels", who, H, W, GBCODEC_MAX_TILE);
    if ((long long)B * K > (1ll << 30)) return fail(GBCODEC_ERR_BAD_SHAPE, "%s: B*K too large", who);
    return GBCODEC_OK;
}

}  // namespace gbc

using namespace gbc;

extern "C" {

int gbcodec_abi_version(void) { return GBCODEC_ABI_VERSION; }

unsigned long long gbcodec_launch_count(void) { return launch_count(); }

const char* gbcodec_status_string(int status) {
    switch (status) {
        case GBCODEC_OK: return "ok";
        case GBCODEC_ERR_NULL_POINTER: return "null pointer";
        case GBCODEC_ERR_BAD_SHAPE: return "bad shape";
        case GBCODEC_ERR_UNALIGNED: return "unaligned pointer";
        case GBCODEC_ERR_BAD_ARGUMENT: return "bad argument";
        case GBCODEC_ERR_WORKSPACE: return "workspace missing or too small";
        case GBCODEC_ERR_CUDA: return "CUDA error";
        case GBCODEC_ERR_PEER_TIMEOUT: return "peer exchange timed out";
        default: return "unknown status";
    }
}

const char* gbcodec_last_error(void) { return g_last_error; }

int gbcodec_encode_mode_f32(const float* d_kps, const float* d_vis, float* d_target, float* d_weight,
                            int B, int K, int H, int W, float in_w, float in_h, double sigma, int mode, void* stream) {
    int st = check_tile_shape("encode", B, K, H, W);
    if (st) return st;
    if (!d_kps || !d_vis || !d_target || !d_weight) return fail(GBCODEC_ERR_NULL_POINTER, "encode: NULL pointer");
    if (!aligned16(d_target)) return fail(GBCODEC_ERR_UNALIGNED, "encode: d_target must be 16-byte aligned");
    if (!(sigma > 0.0) || !(in_w > 0.f) || !(in_h > 0.f)) return fail(GBCODEC_ERR_BAD_ARGUMENT, "encode: sigma and input size must be positive");
    return launch_encode(d_kps, d_vis, d_target, d_weight, B, K, H, W, in_w, in_h, sigma, mode, (cudaStream_t)stream);
}

int gbcodec_encode_f32(const float* d_kps, const float* d_vis, float* d_target, float* d_weight,
                       int B, int K, int H, int W, float in_w, float in_h, double sigma, void* stream) {
    return gbcodec_encode_mode_f32(d_kps, d_vis, d_target, d_weight, B, K, H, W, in_w, in_h, sigma, GBCODEC_ENCODE_PATCH, stream);
}

int gbcodec_decode_f32(const float* d_hm, const float* d_hm_flipped, const int32_t* d_flip_perm,
                       const float* d_off, const float* d_alpha_param, const float* d_fusion_weight,
                       int B, int K, int H, int W, int local_radius, unsigned flags,
                       float* d_coords, float* d_scores, int32_t* d_centre, void* stream) {
    int st = check_tile_shape("decode", B, K, H, W);
    if (st) return st;
    if (!d_hm || !d_coords || !d_scores) return fail(GBCODEC_ERR_NULL_POINTER, "decode: NULL pointer");
    if (flags & ~7u) return fail(GBCODEC_ERR_BAD_ARGUMENT, "decode: unknown flags 0x%x", flags);
    if ((flags & GBCODEC_DECODE_REFINE) && !d_alpha_param) return fail(GBCODEC_ERR_NULL_POINTER, "decode: d_alpha_param is NULL");
    if ((flags & GBCODEC_DECODE_APPLY_OFFSET) && (!d_off || !d_fusion_weight)) return fail(GBCODEC_ERR_NULL_POINTER, "decode: offsets / fusion weight missing");
    if (local_radius < 0 || local_radius > 8) return fail(GBCODEC_ERR_BAD_ARGUMENT, "decode: local_radius=%d", local_radius);
    if (!aligned16(d_hm) || (d_hm_flipped && !aligned16(d_hm_flipped))) return fail(GBCODEC_ERR_UNALIGNED, "decode: heatmaps must be 16-byte aligned");
    return launch_decode(d_hm, d_hm_flipped, d_flip_perm, d_off, d_alpha_param, d_fusion_weight, B, K, H, W,
                         local_radius, flags, d_coords, d_scores, d_centre, (cudaStream_t)stream);
}

int gbcodec_decode_argmax_f32(const float* d_hm, int B, int K, int H, int W, int mode,
                              float* d_coords, float* d_maxvals, int32_t* d_index, void* stream) {
    int st = check_tile_shape("decode_argmax", B, K, H, W);
    if (st) return st;
    if (!d_hm || !d_coords || !d_maxvals) return fail(GBCODEC_ERR_NULL_POINTER, "decode_argmax: NULL pointer");
    if (mode < GBCODEC_ARGMAX_PLAIN || mode > GBCODEC_ARGMAX_TAYLOR) return fail(GBCODEC_ERR_BAD_ARGUMENT, "decode_argmax: mode=%d", mode);
    if (!aligned16(d_hm)) return fail(GBCODEC_ERR_UNALIGNED, "decode_argmax: d_hm must be 16-byte aligned");
    return launch_argmax(d_hm, B, K, H, W, mode, d_coords, d_maxvals, d_index, (cudaStream_t)stream);
}

int gbcodec_refine_centroid_f32(const float* d_hm, const float* d_coords_in, int B, int K, int H, int W,
                                int window, float* d_coords_out, void* stream) {
    int st = check_tile_shape("refine_centroid", B, K, H, W);
    if (st) return st;
    if (!d_hm || !d_coords_in || !d_coords_out) return fail(GBCODEC_ERR_NULL_POINTER, "refine_centroid: NULL pointer");
    if (window < 1 || window > 31) return fail(GBCODEC_ERR_BAD_ARGUMENT, "refine_centroid: window=%d", window);
    return launch_centroid(d_hm, d_coords_in, B, K, H, W, window, d_coords_out, (cudaStream_t)stream);
}

int gbcodec_postprocess_f32(const gbcodec_postprocess_desc* desc, const float* d_hm, const float* d_regression,
                            const float* d_center, const float* d_scale,
                            float* d_preds, float* d_maxvals, float* d_mask, void* d_workspace16, void* stream) {
    if (!desc) return fail(GBCODEC_ERR_NULL_POINTER, "postprocess: desc is NULL");
    int st = check_tile_shape("postprocess", desc->B, desc->K, desc->H, desc->W);
    if (st) return st;
    if (!d_hm || !d_preds || !d_maxvals) return fail(GBCODEC_ERR_NULL_POINTER, "postprocess: NULL pointer");
    if (!aligned16(d_hm)) return fail(GBCODEC_ERR_UNALIGNED, "postprocess: d_hm must be 16-byte aligned");
    if (desc->argmax_mode < GBCODEC_ARGMAX_PLAIN || desc->argmax_mode > GBCODEC_ARGMAX_TAYLOR) return fail(GBCODEC_ERR_BAD_ARGUMENT, "postprocess: argmax_mode=%d", desc->argmax_mode);
    if (desc->refine_window < 0 || desc->refine_window > 31) return fail(GBCODEC_ERR_BAD_ARGUMENT, "postprocess: refine_window=%d", desc->refine_window);
    if (desc->transform && (!d_center || !d_scale)) return fail(GBCODEC_ERR_NULL_POINTER, "postprocess: transform needs d_center and d_scale");
    if (desc->transform && (!(desc->input_w > 0.f) || !(desc->input_h > 0.f))) return fail(GBCODEC_ERR_BAD_ARGUMENT, "postprocess: input size must be positive");
    if ((desc->scale_to_image || d_regression) && !(desc->image_size > 0.f)) return fail(GBCODEC_ERR_BAD_ARGUMENT, "postprocess: image_size must be positive");
    if (d_regression && !d_workspace16) return fail(GBCODEC_ERR_WORKSPACE, "postprocess: 16-byte workspace needed with d_regression");
    return launch_postprocess(desc, d_hm, d_regression, d_center, d_scale, d_preds, d_maxvals, d_mask, d_workspace16, (cudaStream_t)stream);
}

int gbcodec_coords_to_image_f32(const float* d_coords_in, const float* d_center, const float* d_scale,
                                int B, int K, int H, int W, float in_w, float in_h, float* d_coords_out, void* stream) {
    if (B <= 0 || K <= 0 || H <= 0 || W <= 0) return fail(GBCODEC_ERR_BAD_SHAPE, "coords_to_image: B,K,H,W must be positive");
    if (!d_coords_in || !d_center || !d_scale || !d_coords_out) return fail(GBCODEC_ERR_NULL_POINTER, "coords_to_image: NULL pointer");
    if (!(in_w > 0.f) || !(in_h > 0.f)) return fail(GBCODEC_ERR_BAD_ARGUMENT, "coords_to_image: input size must be positive");
    return launch_coords_to_image(d_coords_in, d_center, d_scale, B, K, H, W, in_w, in_h, d_coords_out, (cudaStream_t)stream);
}

size_t gbcodec_combined_workspace_bytes(int B, int K) {
    if (B <= 0 || K <= 0) return 0;
    return combined_workspace_bytes(B, K);
}

int gbcodec_combined_loss_f32(const gbcodec_combined_desc* desc,
                              const float* d_pred, const float* d_target, const float* d_weight,
                              const float* d_coords, const float* d_refined, const float* d_target_coords,
                              const float* d_grad_scale, float* d_losses5,
                              float* d_grad_pred, float* d_grad_coords, float* d_grad_refined,
                              void* d_workspace, size_t workspace_bytes, void* stream) {
    return combined_loss(desc, d_pred, d_target, d_weight, d_coords, d_refined, d_target_coords, d_grad_scale, d_losses5,
                         d_grad_pred, d_grad_coords, d_grad_refined, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

int gbcodec_combined_loss_f16(const gbcodec_combined_desc* desc,
                              const void* d_pred_f16, const float* d_target, const float* d_weight,
                              const float* d_coords, const float* d_refined, const float* d_target_coords,
                              const float* d_grad_scale, float* d_losses5,
                              void* d_grad_pred_f16, float* d_grad_coords, float* d_grad_refined,
                              void* d_workspace, size_t workspace_bytes, void* stream) {
    return combined_loss(desc, reinterpret_cast<const float*>(d_pred_f16), d_target, d_weight, d_coords, d_refined, d_target_coords,
                         d_grad_scale, d_losses5, reinterpret_cast<float*>(d_grad_pred_f16), d_grad_coords, d_grad_refined,
                         d_workspace, workspace_bytes, (cudaStream_t)stream, 1);
}

int gbcodec_combined_loss_backward_f16(const gbcodec_combined_desc* desc,
                              const void* d_pred_f16, const float* d_target, const float* d_weight,
                              const float* d_coords, const float* d_refined, const float* d_target_coords,
                              const float* d_grad_scale, const float* d_grad_losses5,
                              void* d_grad_pred_f16, float* d_grad_coords, float* d_grad_refined,
                              void* d_workspace, size_t workspace_bytes, void* stream) {
    return combined_loss_backward(desc, reinterpret_cast<const float*>(d_pred_f16), d_target, d_weight, d_coords, d_refined,
                                  d_target_coords, d_grad_scale, d_grad_losses5, reinterpret_cast<float*>(d_grad_pred_f16),
                                  d_grad_coords, d_grad_refined, d_workspace, workspace_bytes, (cudaStream_t)stream, 1);
}

int gbcodec_combined_loss_backward_f32(const gbcodec_combined_desc* desc,
                              const float* d_pred, const float* d_target, const float* d_weight,
                              const float* d_coords, const float* d_refined, const float* d_target_coords,
                              const float* d_grad_scale, const float* d_grad_losses5,
                              float* d_grad_pred, float* d_grad_coords, float* d_grad_refined,
                              void* d_workspace, size_t workspace_bytes, void* stream) {
    return combined_loss_backward(desc, d_pred, d_target, d_weight, d_coords, d_refined, d_target_coords, d_grad_scale,
                                  d_grad_losses5, d_grad_pred, d_grad_coords, d_grad_refined, d_workspace, workspace_bytes,
                                  (cudaStream_t)stream);
}

int gbcodec_heatmap_step_f32(const float* d_hm, const float* d_target, const float* d_weight, const float* d_gt_kps,
                             int B, int K, int H, int W, float in_w, float in_h, double sigma,
                             int use_target_weight, int norm_batch, const float* d_grad_scale,
                             float* d_loss, float* d_grad_hm, int argmax_mode,
                             float* d_coords, float* d_maxvals, int32_t* d_index,
                             void* d_workspace, size_t workspace_bytes, void* stream) {
    int st = check_tile_shape("heatmap_step", B, K, H, W);
    if (st) return st;
    if (argmax_mode < GBCODEC_ARGMAX_PLAIN || argmax_mode > GBCODEC_ARGMAX_TAYLOR) return fail(GBCODEC_ERR_BAD_ARGUMENT, "heatmap_step: argmax_mode=%d", argmax_mode);
    if (!d_target && (!(in_w > 0.f) || !(in_h > 0.f))) return fail(GBCODEC_ERR_BAD_ARGUMENT, "heatmap_step: input size must be positive");
    return heatmap_step(d_hm, d_target, d_weight, d_gt_kps, B, K, H, W, in_w, in_h, sigma, use_target_weight, norm_batch, d_grad_scale,
                        d_loss, d_grad_hm, argmax_mode, d_coords, d_maxvals, d_index, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t gbcodec_loss_workspace_bytes(int B, int K, int H, int W) {
    (void)H; (void)W;
    if (B <= 0 || K <= 0) return 0;
    return loss_workspace_bytes(B, K);
}

int gbcodec_loss_denominators_f32(const gbcodec_loss_desc* desc, const float* d_weight, const float* d_gt_kps,
                                  int target_given, float* d_out2_raw_sums,
                                  void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!desc) return fail(GBCODEC_ERR_NULL_POINTER, "denominators: desc is NULL");
    return loss_denominators(desc, d_weight, d_gt_kps, target_given, d_out2_raw_sums, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

int gbcodec_fusion_loss_f32(const gbcodec_loss_desc* desc,
                            const float* d_hm, const float* d_off, const float* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps,
                            const float* d_denoms, const float* d_grad_scale,
                            float* d_losses7, float* d_grad_hm, float* d_grad_off, float* d_grad_var,
                            void* d_workspace, size_t workspace_bytes, void* stream) {
    return fusion_loss(desc, d_hm, d_off, d_var, d_target, d_weight, d_gt_kps, d_denoms, d_grad_scale,
                       d_losses7, d_grad_hm, d_grad_off, d_grad_var,
                       nullptr, nullptr, 0, 0u, nullptr, nullptr, d_workspace, workspace_bytes, (cudaStream_t)stream, nullptr, nullptr, 0, nullptr, nullptr);
}

int gbcodec_fusion_step_f32(const gbcodec_loss_desc* desc,
                            const float* d_hm, const float* d_off, const float* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps,
                            const float* d_denoms, const float* d_grad_scale,
                            float* d_losses7, float* d_grad_hm, float* d_grad_off, float* d_grad_var,
                            const float* d_alpha_param, const float* d_fusion_weight, int local_radius, unsigned decode_flags,
                            float* d_coords, float* d_scores,
                            void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!d_coords || !d_scores) return fail(GBCODEC_ERR_NULL_POINTER, "step: d_coords / d_scores is NULL");
    return fusion_loss(desc, d_hm, d_off, d_var, d_target, d_weight, d_gt_kps, d_denoms, d_grad_scale,
                       d_losses7, d_grad_hm, d_grad_off, d_grad_var,
                       d_alpha_param, d_fusion_weight, local_radius, decode_flags, d_coords, d_scores,
                       d_workspace, workspace_bytes, (cudaStream_t)stream, nullptr, nullptr, 0, nullptr, nullptr);
}

int gbcodec_fusion_step_sharded_f32(const gbcodec_loss_desc* desc,
                            const float* d_hm, const float* d_off, const float* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps, const float* d_grad_scale,
                            float* d_losses7, float* d_grad_hm, float* d_grad_off, float* d_grad_var,
                            const float* d_alpha_param, const float* d_fusion_weight, int local_radius, unsigned decode_flags,
                            float* d_coords, float* d_scores, float* d_denoms_out,
                            const float* d_denoms_global, int defer_losses,
                            void* d_workspace, size_t workspace_bytes, void* peer_ctx, void* stream) {
    if (!peer_ctx) return fail(GBCODEC_ERR_NULL_POINTER, "sharded step: peer_ctx is NULL");
    if ((d_coords == nullptr) != (d_scores == nullptr)) return fail(GBCODEC_ERR_NULL_POINTER, "sharded step: give d_coords and d_scores or neither");
    return fusion_loss(desc, d_hm, d_off, d_var, d_target, d_weight, d_gt_kps, d_denoms_global, d_grad_scale,
                       d_losses7, d_grad_hm, d_grad_off, d_grad_var,
                       d_alpha_param, d_fusion_weight, local_radius, decode_flags, d_coords, d_scores,
                       d_workspace, workspace_bytes, (cudaStream_t)stream, peer_ctx, d_denoms_out, 0, nullptr, nullptr, defer_losses ? 1 : 0);
}

int gbcodec_peer_denominators_f32(const gbcodec_loss_desc* desc, const float* d_weight, const float* d_gt_kps, int target_given,
                                  float* d_out2_global_sums, void* d_workspace, size_t workspace_bytes, void* peer_ctx, void* stream) {
    return peer_denominators(desc, d_weight, d_gt_kps, target_given, d_out2_global_sums, d_workspace, workspace_bytes, peer_ctx, (cudaStream_t)stream);
}

int gbcodec_peer_collect_losses_f32(void* peer_ctx, int steps_back, float* d_losses7, void* stream) {
    return peer_collect_losses(peer_ctx, steps_back, d_losses7, (cudaStream_t)stream);
}

int gbcodec_peer_create(int rank, int world, void** ctx_out, unsigned char* handle_out) { return peer_create(rank, world, ctx_out, handle_out); }
int gbcodec_peer_connect(void* ctx, const unsigned char* all_handles) { return peer_connect(ctx, all_handles); }
int gbcodec_peer_status(void* ctx, int* h_timeouts) { return peer_status(ctx, h_timeouts); }
int gbcodec_peer_destroy(void* ctx) { return peer_destroy(ctx); }
int gbcodec_peer_set_timeout(void* ctx, double seconds) { return peer_set_timeout(ctx, seconds); }

int gbcodec_fusion_step_vmean_f32(const gbcodec_loss_desc* desc,
                            const float* d_hm, const float* d_off, const float* d_var_mean, const float* d_target,
                            const float* d_weight, const float* d_gt_kps,
                            const float* d_denoms, const float* d_grad_scale,
                            float* d_losses7, float* d_grad_hm, float* d_grad_off, float* d_grad_var_mean,
                            const float* d_alpha_param, const float* d_fusion_weight, int local_radius, unsigned decode_flags,
                            float* d_coords, float* d_scores,
                            void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!d_var_mean) return fail(GBCODEC_ERR_NULL_POINTER, "step_vmean: d_var_mean is NULL");
    if ((d_coords == nullptr) != (d_scores == nullptr)) return fail(GBCODEC_ERR_NULL_POINTER, "step_vmean: give d_coords and d_scores or neither");
    return fusion_loss(desc, d_hm, d_off, nullptr, d_target, d_weight, d_gt_kps, d_denoms, d_grad_scale,
                       d_losses7, d_grad_hm, d_grad_off, nullptr,
                       d_alpha_param, d_fusion_weight, local_radius, decode_flags, d_coords, d_scores,
                       d_workspace, workspace_bytes, (cudaStream_t)stream, nullptr, nullptr, 0, d_var_mean, d_grad_var_mean);
}

int gbcodec_fusion_loss_backward_f32(const gbcodec_loss_desc* desc,
                            const float* d_hm, const float* d_off, const float* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps,
                            const float* d_denoms, const float* d_grad_scale, const float* d_grad_losses7,
                            float* d_grad_hm, float* d_grad_off, float* d_grad_var,
                            float* d_held6, int held_valid, int workspace_from_forward,
                            void* d_workspace, size_t workspace_bytes, void* stream) {
    return fusion_loss_backward(desc, d_hm, d_off, d_var, d_target, d_weight, d_gt_kps, d_denoms, d_grad_scale,
                                d_grad_losses7, d_grad_hm, d_grad_off, d_grad_var, d_workspace, workspace_bytes,
                                (cudaStream_t)stream, 0, 1, d_held6, held_valid, workspace_from_forward);
}

int gbcodec_fusion_step_f16(const gbcodec_loss_desc* desc,
                            const void* d_hm, const void* d_off, const void* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps, const float* d_denoms, const float* d_grad_scale,
                            float* d_losses7, void* d_grad_hm, void* d_grad_off, void* d_grad_var,
                            const float* d_alpha_param, const float* d_fusion_weight, int local_radius, unsigned decode_flags,
                            float* d_coords, float* d_scores, void* d_workspace, size_t workspace_bytes, void* stream) {
    if ((d_coords == nullptr) != (d_scores == nullptr)) return fail(GBCODEC_ERR_NULL_POINTER, "step_f16: give d_coords and d_scores or neither");
    return fusion_loss(desc, (const float*)d_hm, (const float*)d_off, (const float*)d_var, d_target, d_weight, d_gt_kps, d_denoms, d_grad_scale,
                       d_losses7, (float*)d_grad_hm, (float*)d_grad_off, (float*)d_grad_var,
                       d_alpha_param, d_fusion_weight, local_radius, decode_flags, d_coords, d_scores,
                       d_workspace, workspace_bytes, (cudaStream_t)stream, nullptr, nullptr, 1, nullptr, nullptr);
}

int gbcodec_fusion_loss_backward_f16(const gbcodec_loss_desc* desc,
                            const void* d_hm, const void* d_off, const void* d_var, const float* d_target,
                            const float* d_weight, const float* d_gt_kps, const float* d_denoms,
                            const float* d_grad_scale, int gradients_stored, const float* d_grad_losses7,
                            void* d_grad_hm, void* d_grad_off, void* d_grad_var,
                            float* d_held6, int held_valid, int workspace_from_forward,
                            void* d_workspace, size_t workspace_bytes, void* stream) {
    return fusion_loss_backward(desc, (const float*)d_hm, (const float*)d_off, (const float*)d_var, d_target, d_weight, d_gt_kps, d_denoms,
                                d_grad_scale, d_grad_losses7, (float*)d_grad_hm, (float*)d_grad_off, (float*)d_grad_var,
                                d_workspace, workspace_bytes, (cudaStream_t)stream, 1, gradients_stored ? 1 : 0,
                                d_held6, held_valid, workspace_from_forward);
}

int gbcodec_softplus_mean_f32(const float* d_raw, float* d_mean, int B, int K, int H, int W, void* stream) {
    int st = check_tile_shape("softplus_mean", B, K, H, W);
    if (st) return st;
    if (!d_raw || !d_mean) return fail(GBCODEC_ERR_NULL_POINTER, "softplus_mean: NULL pointer");
    if (!aligned16(d_raw)) return fail(GBCODEC_ERR_UNALIGNED, "softplus_mean: d_raw must be 16-byte aligned");
    return launch_softplus_mean(d_raw, d_mean, B, K, H, W, (cudaStream_t)stream);
}

int gbcodec_softplus_mean_backward_f32(const float* d_raw, const float* d_grad_mean, float* d_grad_raw, int B, int K, int H, int W, void* stream) {
    int st = check_tile_shape("softplus_mean_backward", B, K, H, W);
    if (st) return st;
    if (!d_raw || !d_grad_mean || !d_grad_raw) return fail(GBCODEC_ERR_NULL_POINTER, "softplus_mean_backward: NULL pointer");
    if (!aligned16(d_raw) || !aligned16(d_grad_raw)) return fail(GBCODEC_ERR_UNALIGNED, "softplus_mean_backward: maps must be 16-byte aligned");
    return launch_softplus_mean_backward(d_raw, d_grad_mean, d_grad_raw, B, K, H, W, (cudaStream_t)stream);
}

int gbcodec_profile_loss_kernel(void* start_event, void* stop_event) {
    if ((start_event == nullptr) != (stop_event == nullptr)) return fail(GBCODEC_ERR_NULL_POINTER, "profile: give both events or neither");
    set_profile_events((cudaEvent_t)start_event, (cudaEvent_t)stop_event);
    return GBCODEC_OK;
}

}  // extern "C"
