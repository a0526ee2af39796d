// TEST INFRASTRUCTURE ONLY.  Host build of the __host__ __device__ camera functions of lichtfeld-studio_b200/csrc/cameras.cuh
// (the same source the projection and ray kernels compile), exported over plain arrays so that tests/test_host_cameras.py
// can check them on a machine without a GPU.  Built by the test itself:
//     nvcc -O2 -shared -Xcompiler -fPIC -o tests/_build/libhost_cameras.so tests/host_cameras.cu
#include "../lichtfeld-studio_b200/csrc/cameras.cuh"

using namespace lfs;

namespace {
CamModel model_of(const float* vm0, const float* vm1, const float* K, int w, int h, int model, int shutter, const float* radial,
                  const float* tangential, const float* prism) {
    return make_cam_model(vm0, vm1, K, (uint32_t)w, (uint32_t)h, model, shutter, radial, tangential, prism);
}
} // namespace

#define CAM_ARGS                                                                                                                 \
    const float *vm0, const float *vm1, const float *K, int w, int h, int model, int shutter, const float *radial,               \
        const float *tangential, const float *prism
#define CAM_PASS vm0, vm1, K, w, h, model, shutter, radial, tangential, prism

extern "C" {

// fisheye set-up: out[0] = max angle, out[1] = slope of the crude inverse
void hc_fisheye_limits(CAM_ARGS, float* out) {
    const CamModel c = model_of(CAM_PASS);
    out[0] = c.fish_max_angle, out[1] = c.fish_back1;
}

// camera-space points [n,3] -> uv [n,2], valid [n]
void hc_cam_project(CAM_ARGS, int n, const float* pc, float margin, float* uv, int* valid) {
    const CamModel c = model_of(CAM_PASS);
    for (int i = 0; i < n; ++i)
        valid[i] = cam_project(c, f3{pc[3 * i], pc[3 * i + 1], pc[3 * i + 2]}, margin, uv[2 * i], uv[2 * i + 1]) ? 1 : 0;
}

// uv [n,2] -> unit camera rays [n,3], valid [n]
void hc_cam_unproject(CAM_ARGS, int n, const float* uv, float* rays, int* valid) {
    const CamModel c = model_of(CAM_PASS);
    for (int i = 0; i < n; ++i) {
        f3 r;
        valid[i] = cam_unproject(c, uv[2 * i], uv[2 * i + 1], r) ? 1 : 0;
        rays[3 * i] = r.x, rays[3 * i + 1] = r.y, rays[3 * i + 2] = r.z;
    }
}

// world points [n,3] -> uv [n,2], valid [n] through the (rolling) shutter
void hc_world_to_image(CAM_ARGS, int n, const float* pw, float margin, float* uv, int* valid) {
    const CamModel c = model_of(CAM_PASS);
    for (int i = 0; i < n; ++i)
        valid[i] = world_to_image(c, f3{pw[3 * i], pw[3 * i + 1], pw[3 * i + 2]}, margin, uv[2 * i], uv[2 * i + 1]) ? 1 : 0;
}

// uv [n,2] -> world ray origin [n,3] / direction [n,3], valid [n]
void hc_pixel_to_world_ray(CAM_ARGS, int n, const float* uv, float* org, float* dir, int* valid) {
    const CamModel c = model_of(CAM_PASS);
    for (int i = 0; i < n; ++i) {
        f3 o, d;
        valid[i] = pixel_to_world_ray(c, uv[2 * i], uv[2 * i + 1], o, d) ? 1 : 0;
        org[3 * i] = o.x, org[3 * i + 1] = o.y, org[3 * i + 2] = o.z;
        dir[3 * i] = d.x, dir[3 * i + 1] = d.y, dir[3 * i + 2] = d.z;
    }
}

// uv [n,2] -> relative frame time [n]
void hc_shutter_time(CAM_ARGS, int n, const float* uv, float* t) {
    const CamModel c = model_of(CAM_PASS);
    for (int i = 0; i < n; ++i) t[i] = shutter_time(c, uv[2 * i], uv[2 * i + 1]);
}

// pose at relative time t [n] -> quaternion (w,x,y,z) [n,4], translation [n,3]; exact-sin variant (the only one on the host)
void hc_shutter_pose(CAM_ARGS, int n, const float* t, float* q, float* tr) {
    const CamModel c = model_of(CAM_PASS);
    for (int i = 0; i < n; ++i) {
        quat4 qq;
        f3 tt;
        shutter_pose<false>(c, t[i], qq, tt);
        q[4 * i] = qq.w, q[4 * i + 1] = qq.x, q[4 * i + 2] = qq.y, q[4 * i + 3] = qq.z;
        tr[3 * i] = tt.x, tr[3 * i + 1] = tt.y, tr[3 * i + 2] = tt.z;
    }
}

// quaternion (w,x,y,z) of a row-major [4,4] pose and its rotation matrix back
void hc_quat_roundtrip(const float* vm, float* q, float* R) {
    const quat4 qq = quat_from_rowmajor_rot(vm);
    q[0] = qq.w, q[1] = qq.x, q[2] = qq.y, q[3] = qq.z;
    quat_to_mat3(qq, R);
}
}
