// Pre-render physics filter on the GPU: collision / support / stability of every candidate pose of the
// movable object against the static scene, all poses of the grid in one launch.
//
// Replaces the serial pybullet loop of reference vision_3d/physics_utils.py:305-372 (per pose: set_pose + up to six
// pairwise_collision queries; 2 211 840 poses for the shelf demo, configs/shelf_demo.json:28).  The control flow is the
// reference's; the collision primitive is an occupancy overlap between the two NeRFs the path already holds -- the occupied
// cell centres of the movable object's density grid, moved by pose . init_pose^-1, against the background model's occupancy
// bitfield (the lookup of the renderer's DDA, nerf_device.cuh:430-447) -- because the reference's primitive lives in pybullet
// on Poisson meshes that do not exist on this path (the test suite states the definition independently in numpy; parity with
// pybullet itself is unpinned).
#include "d2r_march.cuh"

namespace d2r {

struct PhysParams {
    const uint8_t* bits_lin;
    int max_cascade;
    float scale, off[3];
    const float4* pts;          // occupied cell centres of the movable object, world frame, at its initial pose
    int n_pts;
    const float* rel;           // [N,12]: pose . init_pose^-1, rows of the 3x4
    const float* pose_z;        // [N]: z of the candidate pose itself (the below-table rule, physics_utils.py:333-335)
    const uint8_t* valid_in;
    uint8_t* valid_out;
    int N;
    float scene_z, lower, p_dist;
    int stability;
};

// does any point of the object, moved by (R, t) and then by `shift`, land in an occupied cell of the static scene?  one warp
__device__ __forceinline__ bool collide(const PhysParams& P, const float (&R)[12], float sx, float sy, float sz, int lane) {
    for (int base = 0; base < P.n_pts; base += 32) {
        const int i = base + lane;
        bool hit = false;
        if (i < P.n_pts) {
            const float4 p = __ldg(P.pts + i);
            // numpy evaluates every operator separately in float32 (the oracle): no fma contraction here
            float w[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                float v = __fadd_rn(__fadd_rn(__fmul_rn(p.x, R[4 * r + 0]), __fmul_rn(p.y, R[4 * r + 1])), __fmul_rn(p.z, R[4 * r + 2]));
                w[r] = __fadd_rn(v, R[4 * r + 3]);
            }
            w[0] = __fadd_rn(w[0], sx); w[1] = __fadd_rn(w[1], sy); w[2] = __fadd_rn(w[2], sz);
            // NerfDataset::nerf_position_to_ngp (nerf_loader.h:148-151): p * scale + offset, then xyz <- yzx
            const float a = __fadd_rn(__fmul_rn(w[0], P.scale), P.off[0]), b = __fadd_rn(__fmul_rn(w[1], P.scale), P.off[1]),
                        c = __fadd_rn(__fmul_rn(w[2], P.scale), P.off[2]);
            const float qx = b, qy = c, qz = a;
            const uint32_t mip = min(mip_from_pos(qx, qy, qz), (uint32_t)P.max_cascade);
            hit = density_grid_occupied_at(qx, qy, qz, P.bits_lin, mip);
        }
        if (__any_sync(0xffffffffu, hit)) return true;
    }
    return false;
}

__global__ void __launch_bounds__(256) k_phys_check(const __grid_constant__ PhysParams P) {
    const int pose = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (pose >= P.N) return;
    bool valid = P.valid_in[pose] != 0;
    if (valid) {
        float R[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) R[j] = __ldg(P.rel + (size_t)pose * 12 + j);
        // in collision -> invalid (physics_utils.py:317-326)
        if (collide(P, R, 0.f, 0.f, 0.f, lane)) valid = false;
        else if (!(__ldg(P.pose_z + pose) < P.scene_z)) {            // a pose below the table plane counts as supported (:333-335)
            // moved down along gravity it must touch a support (:329-343) ...
            if (!collide(P, R, 0.f, 0.f, -P.lower, lane)) valid = false;
            else if (P.stability) {
                // ... and still touch one when pushed p_dist along +-x, +-y (:351-368)
                const float dx[4] = {P.p_dist, -P.p_dist, 0.f, 0.f}, dy[4] = {0.f, 0.f, P.p_dist, -P.p_dist};
#pragma unroll 1
                for (int k = 0; k < 4 && valid; ++k)
                    if (!collide(P, R, dx[k], dy[k], -P.lower, lane)) valid = false;
            }
        }
    }
    if (lane == 0) P.valid_out[pose] = valid ? 1 : 0;
}

}  // namespace d2r

extern "C" int d2r_phys_check(const d2r_model* bg, const float* fg_points_world_dev, int n_pts, const float* rel_3x4_dev,
                              const float* pose_z_dev, const uint8_t* valid_in_dev, int N, const d2r_phys_cfg* cfg,
                              uint8_t* valid_out_dev, void* stream) {
    using namespace d2r;
    D2R_REQUIRE(bg && fg_points_world_dev && rel_3x4_dev && pose_z_dev && valid_in_dev && cfg && valid_out_dev, "d2r_phys_check: null argument");
    D2R_REQUIRE(N > 0 && n_pts >= 0, "d2r_phys_check: bad sizes");
    DeviceGuard dg(bg->device);
    PhysParams P;
    P.bits_lin = bg->dev.bitfield_lin; P.max_cascade = bg->dev.max_cascade;
    P.scale = cfg->dataset_scale;
    for (int i = 0; i < 3; ++i) P.off[i] = cfg->dataset_offset[i];
    P.pts = (const float4*)fg_points_world_dev; P.n_pts = n_pts;
    P.rel = rel_3x4_dev; P.pose_z = pose_z_dev; P.valid_in = valid_in_dev; P.valid_out = valid_out_dev; P.N = N;
    P.scene_z = cfg->scene_centre_z; P.lower = cfg->unsup_thresh; P.p_dist = cfg->p_dist; P.stability = cfg->stability_check;
    const size_t threads = (size_t)N * 32;
    k_phys_check<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
    count_launch();
    D2R_CUDA(cudaGetLastError());
    return D2R_OK;
}
