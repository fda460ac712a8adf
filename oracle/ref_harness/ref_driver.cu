// ref_driver.cu — ROS-free driver that replays VOLMAPNODE::publishMap's call order
// (reference src/volumetric_mapper.cpp:138-224) on the reference's OWN, unmodified CUDA sources, which
// oracle/build_ref.sh compiles from /root/reference into oracle/_ref/.  TEST INFRASTRUCTURE ONLY: it produces the golden
// fixtures under tests/golden/ (oracle/gen_golden.py) and is the `--impl reference` arm of bench.py.
//
// usage: ref_driver <in.bin> <out.bin|-> [--time] [--halo H]
//   in.bin : written by oracle/ref_io.py (config header + frames)
//   out.bin: per frame  glb_type i8[N], aux i32[N], coc_aux i32[N], pair_dist i32[N], pair_id i32[N], edt f32[N],
//            then a (X+2H)(Y+2H)(Z+2H) box of {alloc i8, type i8, occ u8, pad, dist i32, coc i32[3]} around the volume
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <vector>
#include <string>

#include "map_structure/local_batch.h"
#include "cuda_toolkit/projection.h"
#include "cuda_toolkit/edt/edt_interfaces.h"
#include "par_wave/glb_hash_map.h"
#include "map_structure/pre_map.h"
#include "kernel/point_cloud/pntcld_interfaces.h"
#include "kernel/hokuyo/hokuyo_interfaces.h"
#include "kernel/vlp16/vlp16_interface.h"
#include "kernel/realsense/realsense_interfaces.h"

struct Header {
    int magic, sensor, X, Y, Z, thresh, cutoff_sq, fast, bucket_max, block_max, fmp, r2, nframes;
    int rows, cols, scan_num, ring_num, valid_nan;
    float w, min_h, max_h;
    float theta_inc, theta_min, phi_inc, phi_min, cx, cy, fx, fy;
};

struct BoxVox { signed char alloc, type; unsigned char occ, pad; int dist; int coc[3]; };

__global__ void dump_box(LocMap m, HASH_BASE hb, int H, BoxVox *out)
{
    int bx = m._local_size.x + 2 * H, by = m._local_size.y + 2 * H, bz = m._local_size.z + 2 * H;
    long long n = (long long)bx * by * bz;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int3 l = make_int3((int)(i % bx) - H, (int)((i / bx) % by) - H, (int)(i / ((long long)bx * by)) - H);
    int3 g = m.loc2glb(l);
    BoxVox o; memset(&o, 0, sizeof(o));
    int id = hb.get_alloc_blk_id(get_VB_key(g));
    if (id >= 0) {
        GlbVoxel *v = retrive_vox_D(g, &hb.alloc[id]);
        o.alloc = 1; o.type = v->vox_type; o.occ = v->occ_val; o.dist = v->dist_sq;
        o.coc[0] = v->coc_glb.x; o.coc[1] = v->coc_glb.y; o.coc[2] = v->coc_glb.z;
    }
    out[i] = o;
}

__global__ void split_pair(const Dist_id *p, int n, int *dist, int *id)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dist[i] = p[i].sq_dist[0]; id[i] = p[i].parent_loc_id[1];
}

static void setup_plans(LocMap *lm, cuttHandle *plan)
{
    // VOLMAPNODE::setupRotationPlan, src/volumetric_mapper.cpp:344-373
    int Dx = lm->_local_size.x, Dy = lm->_local_size.y, Dz = lm->_local_size.z;
    if (Dz == 1) {
        int d0[2] = { Dx, Dy }, p0[2] = { 1, 0 }, d1[2] = { Dy, Dx }, p1[2] = { 1, 0 };
        cuttPlan(&plan[0], 2, d0, p0, sizeof(int), nullptr);
        cuttPlan(&plan[1], 2, d1, p1, sizeof(int), nullptr);
    } else {
        int d0[3] = { Dx, Dy, Dz }, p0[3] = { 1, 0, 2 }, d1[3] = { Dy, Dx, Dz }, p1[3] = { 0, 2, 1 }, d2[3] = { Dy, Dz, Dx }, p2[3] = { 2, 0, 1 };
        cuttPlan(&plan[0], 3, d0, p0, sizeof(int), nullptr);
        cuttPlan(&plan[1], 3, d1, p1, sizeof(int), nullptr);
        cuttPlan(&plan[2], 3, d2, p2, sizeof(int), nullptr);
    }
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin|- [--time] [--halo H]\n", argv[0]); return 2; }
    bool timing = false; int H = 4;
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--time")) timing = true;
        else if (!strcmp(argv[i], "--halo") && i + 1 < argc) H = atoi(argv[++i]);
    }
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) { perror("in"); return 2; }
    Header h;
    if (fread(&h, sizeof(h), 1, fi) != 1 || h.magic != 0x47494531) { fprintf(stderr, "bad header\n"); return 2; }
    FILE *fo = nullptr;
    if (strcmp(argv[2], "-")) { fo = fopen(argv[2], "wb"); if (!fo) { perror("out"); return 2; } }

    // VOLMAPNODE ctor, src/volumetric_mapper.cpp:70-126
    int3 sz = make_int3(h.X, h.Y, h.Z);
    LocMap *lm = new LocMap(h.w, sz, (unsigned char)h.thresh, h.min_h, h.max_h, h.cutoff_sq, h.fast != 0);
    lm->create_gpu_map();
    cuttHandle plan[3] = { 0, 0, 0 };
    setup_plans(lm, plan);
    GlbHashMap *hm = new GlbHashMap(lm->_bdr_num, lm->_local_size, h.bucket_max, h.block_max);
    hm->setLocMap(lm);
    Ext_Obs_Wrapper *ext_obs = new Ext_Obs_Wrapper(1);
    warmupCuda();

    const int N = h.X * h.Y * h.Z;
    size_t max_payload = 0;
    std::vector<std::vector<float>> payloads(h.nframes);
    std::vector<float> poses((size_t)h.nframes * 7);
    for (int f = 0; f < h.nframes; f++) {
        int n;
        if (fread(&poses[(size_t)f * 7], 4, 7, fi) != 7 || fread(&n, 4, 1, fi) != 1) { fprintf(stderr, "short file\n"); return 2; }
        payloads[f].resize(n);
        if (n && fread(payloads[f].data(), 4, n, fi) != (size_t)n) { fprintf(stderr, "short payload\n"); return 2; }
        if ((size_t)n > max_payload) max_payload = n;
    }
    fclose(fi);
    float *d_in = nullptr;
    GPU_MALLOC(&d_in, (max_payload + 4) * sizeof(float));

    std::vector<char> buf8(N); std::vector<int> bufi(N); std::vector<float> buff(N);
    int *d_dist = nullptr, *d_id = nullptr; BoxVox *d_box = nullptr;
    long long nbox = (long long)(h.X + 2 * H) * (h.Y + 2 * H) * (h.Z + 2 * H);
    std::vector<BoxVox> hbox;
    if (fo) { GPU_MALLOC(&d_dist, N * 4); GPU_MALLOC(&d_id, N * 4); GPU_MALLOC(&d_box, nbox * sizeof(BoxVox)); hbox.resize(nbox); }

    double tot_ogm = 0, tot_edt = 0;
    for (int f = 0; f < h.nframes; f++) {
        const float *ps = &poses[(size_t)f * 7];
        int time = f + 1;   // _time++, volumetric_mapper.cpp:144
        tf::Transform trans(tf::Quaternion(ps[1], ps[2], ps[3], ps[0]), tf::Vector3(ps[4], ps[5], ps[6]));
        auto t0 = std::chrono::steady_clock::now();
        Projection proj = trans2proj(trans);
        lm->calculate_pivot_origin(proj.origin);
        lm->calculate_update_pivot(proj.origin);
        int3 *keys = thrust::raw_pointer_cast(hm->VB_keys_loc_D.data());
        int n = (int)payloads[f].size();
        if (n) GPU_MEMCPY_H2D(d_in, payloads[f].data(), n * sizeof(float));   // the MapMakers' per-frame H2D copy
        if (h.sensor == 0) {
            PntcldParam pp(n / 3); pp.valid_pnt_count = n / 3;
            PNTCLD_RAYCAST::localOGMKernels(lm, (float3 *)d_in, proj, pp, keys, time, h.fmp != 0, h.r2);
        } else if (h.sensor == 1) {
            ScanParam sp(h.scan_num, 30.f, h.theta_inc, h.theta_min);
            HOKUYO_FAST::localOGMKernels(lm, d_in, proj, sp, keys, h.fmp != 0, h.r2);
        } else if (h.sensor == 2) {
            MulScanParam mp(h.scan_num, h.ring_num, 10.f, h.theta_inc, h.theta_min, h.phi_inc, h.phi_min);
            VLP_FAST::localOGMKernels(lm, d_in, proj, mp, keys, h.fmp != 0, h.r2);
        } else {
            CamParam cp(h.rows, h.cols, h.cx, h.cy, h.fx, h.fy, h.valid_nan != 0);
            REALSENSE_FAST::localOGMKernels(lm, d_in, proj, cp, keys, h.fmp != 0, h.r2);
        }
        hm->updateHashOGM(h.sensor == 0, time, false, ext_obs);
        GPU_DEV_SYNC();
        auto t1 = std::chrono::steady_clock::now();
        EDT_OCC::batchEDTUpdate(lm, plan, time);
        hm->mergeNewObsv(time, false);
        GPU_DEV_SYNC();
        auto t2 = std::chrono::steady_clock::now();
        double ogm_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        double edt_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
        tot_ogm += ogm_ms; tot_edt += edt_ms;
        if (timing) printf("frame %d ogm_ms %.4f edt_ms %.4f\n", f, ogm_ms, edt_ms);
        if (fo) {
            GPU_MEMCPY_D2H(buf8.data(), lm->_glb_type, N); fwrite(buf8.data(), 1, N, fo);
            GPU_MEMCPY_D2H(bufi.data(), lm->_aux, N * 4); fwrite(bufi.data(), 4, N, fo);
            GPU_MEMCPY_D2H(bufi.data(), lm->_coc_idx_aux, N * 4); fwrite(bufi.data(), 4, N, fo);
            split_pair<<<(N + 255) / 256, 256>>>(lm->_dist_id_pair, N, d_dist, d_id);
            GPU_MEMCPY_D2H(bufi.data(), d_dist, N * 4); fwrite(bufi.data(), 4, N, fo);
            GPU_MEMCPY_D2H(bufi.data(), d_id, N * 4); fwrite(bufi.data(), 4, N, fo);
            GPU_MEMCPY_D2H(buff.data(), lm->_edt_D, N * 4); fwrite(buff.data(), 4, N, fo);
            dump_box<<<(unsigned)((nbox + 255) / 256), 256>>>(*lm, *(hm->hash_table_D), H, d_box);
            GPU_MEMCPY_D2H(hbox.data(), d_box, nbox * sizeof(BoxVox)); fwrite(hbox.data(), sizeof(BoxVox), nbox, fo);
        }
    }
    if (fo) fclose(fo);
    printf("total frames %d ogm_ms %.4f edt_ms %.4f\n", h.nframes, tot_ogm, tot_edt);
    return 0;
}
