// A GPU "planner" built together with the mapper (README.md:163-170): integrates a few frames through the C++ surface of
// include/gie_compat, then reads N random global voxels from a kernel through gie_device_view + get_VB_key / get_voxID_in_VB
// and compares every field with the host mirror of the same blocks (GlbHashMap::VB_values_H via export).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>
#include "map_structure/local_batch.h"
#include "par_wave/glb_hash_map.h"
#include "par_wave/gie_device_view.cuh"
#include "kernel/ogm_interfaces.h"
#include "cuda_toolkit/edt/edt_interfaces.h"

__global__ void k_query(gie_device_view v, const int3 *q, int n, GlbVoxel *out, int *found)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    found[i] = gie_dv_voxel(v, q[i], &out[i]) ? 1 : 0;
}

int main()
{
    const int3 size = make_int3(64, 64, 32);
    LocMap lm(0.1f, size, 180, -10.f, 10.f, 64, false);
    lm.create_gpu_map();
    GlbHashMap hm(lm._bdr_num, lm._local_size, 4000, 12000);
    hm.setLocMap(&lm);
    cuttHandle plan[3] = { 0, 0, 0 };
    std::vector<float3> pts(3000);
    float3 *d_pts = nullptr;
    cudaMalloc(&d_pts, pts.size() * sizeof(float3));
    for (int f = 0; f < 4; f++) {
        Projection proj;
        const float q[4] = { 1.f, 0.f, 0.f, 0.f }, t[3] = { 0.2f * f, -0.1f, 1.5f };
        gie_make_projection(q, t, proj.L2G.data, proj.G2L.data);
        proj.origin = make_float3(t[0], t[1], t[2]);
        for (size_t i = 0; i < pts.size(); i++) pts[i] = make_float3(-2.5f + 0.0017f * i, 1.2f + 0.3f * ((i * 7) % 5), -0.9f + 0.001f * (i % 900));
        cudaMemcpy(d_pts, pts.data(), pts.size() * sizeof(float3), cudaMemcpyHostToDevice);
        lm.calculate_pivot_origin(proj.origin); lm.calculate_update_pivot(proj.origin);
        PntcldParam pp((int)pts.size()); pp.valid_pnt_count = (int)pts.size();
        PNTCLD_RAYCAST::localOGMKernels(&lm, d_pts, proj, pp, thrust::raw_pointer_cast(hm.VB_keys_loc_D.data()), f + 1, false, 0);
        hm.updateHashOGM(true, f + 1, false, nullptr);
        EDT_OCC::batchEDTUpdate(&lm, plan, f + 1);
        hm.mergeNewObsv(f + 1, false);
    }
    // host mirror of every block
    int nb = 0;
    GIE_CHECK(gie_hashmap_num_blocks(hm.handle(), &nb));
    std::vector<int32_t> keys(3 * (size_t)nb);
    std::vector<gie_glbvoxel> vox(512 * (size_t)nb);
    GIE_CHECK(gie_hashmap_export_blocks(hm.handle(), keys.data(), vox.data(), nb));
    std::map<std::tuple<int, int, int>, int> lut;
    for (int b = 0; b < nb; b++) lut[{ keys[3 * b], keys[3 * b + 1], keys[3 * b + 2] }] = b;
    // queries: voxels of allocated blocks and voxels around them
    const int n = 20000;
    std::vector<int3> q(n);
    srand(7);
    for (int i = 0; i < n; i++) {
        int b = rand() % nb;
        q[i] = make_int3(keys[3 * b] * 8 + rand() % 24 - 8, keys[3 * b + 1] * 8 + rand() % 24 - 8, keys[3 * b + 2] * 8 + rand() % 24 - 8);
    }
    int3 *d_q; GlbVoxel *d_out; int *d_found;
    cudaMalloc(&d_q, n * sizeof(int3)); cudaMalloc(&d_out, n * sizeof(GlbVoxel)); cudaMalloc(&d_found, n * sizeof(int));
    cudaMemcpy(d_q, q.data(), n * sizeof(int3), cudaMemcpyHostToDevice);
    gie_device_view view;
    GIE_CHECK(gie_hashmap_device_view(hm.handle(), &view));
    GIE_CHECK(gie_sync(hm.handle()));
    k_query<<<(n + 255) / 256, 256>>>(view, d_q, n, d_out, d_found);
    std::vector<GlbVoxel> out(n);
    std::vector<int> found(n);
    cudaMemcpy(out.data(), d_out, n * sizeof(GlbVoxel), cudaMemcpyDeviceToHost);
    if (cudaMemcpy(found.data(), d_found, n * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 1; }
    int hits = 0, bad = 0;
    for (int i = 0; i < n; i++) {
        auto it = lut.find({ get_VB_key(q[i]).x, get_VB_key(q[i]).y, get_VB_key(q[i]).z });
        if ((it != lut.end()) != (found[i] != 0)) { bad++; continue; }
        if (!found[i]) continue;
        hits++;
        const gie_glbvoxel &h = vox[(size_t)it->second * 512 + get_voxID_in_VB(q[i])];
        if (memcmp(&h.dist_sq, &out[i].dist_sq, 4) || h.vox_type != out[i].vox_type || h.occ_val != out[i].occ_val ||
            h.coc_glb[0] != out[i].coc_glb.x || h.coc_glb[1] != out[i].coc_glb.y || h.coc_glb[2] != out[i].coc_glb.z ||
            h.dist_id_pair != out[i].dist_id_pair.ulong || h.update_ct != out[i].update_ct || h.wave_layer != out[i].wave_layer) bad++;
    }
    printf("device view: %d queries, %d in allocated blocks, %d mismatches\n", n, hits, bad);
    if (bad || hits < n / 20) return 1;
    printf("device view OK\n");
    return 0;
}
