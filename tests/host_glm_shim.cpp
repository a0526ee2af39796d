// TEST INFRASTRUCTURE ONLY.  Exposes the symbols of oracle/glm_shim (the stand-in for glm under which the UNMODIFIED reference
// gsplat kernels are compiled into oracle/_ref/libgsplat_ref.so) over plain arrays, so that tests/test_glm_shim.py can hold each
// of them to glm's documented conventions with numpy.  Matrices cross this boundary as ROW-MAJOR numpy arrays A[r][c]; the
// glm object is built / read through m[c][r], which is exactly the convention under test.
//     g++ -std=c++17 -O2 -ffp-contract=off -shared -fPIC -Ioracle/glm_shim -o tests/_build/libhost_glm_shim.so tests/host_glm_shim.cpp
#include <math.h> // global float overloads of sqrt / sin / acos, as under nvcc (the shim calls them unqualified)

#include <glm/glm.hpp>
#include <glm/gtc/quaternion.hpp>
#include <glm/gtc/type_ptr.hpp>
#include <glm/gtx/matrix_operation.hpp>
#include <glm/gtx/quaternion.hpp>

namespace {
template <int C, int R>
glm::mat<C, R, float> load(const float* a /* row-major [R][C] */) {
    glm::mat<C, R, float> m;
    for (int c = 0; c < C; ++c)
        for (int r = 0; r < R; ++r) m[c][r] = a[r * C + c];
    return m;
}
template <int C, int R>
void store(const glm::mat<C, R, float>& m, float* a /* row-major [R][C] */) {
    for (int c = 0; c < C; ++c)
        for (int r = 0; r < R; ++r) a[r * C + c] = m[c][r];
}
glm::quat loadq(const float* q /* w x y z */) { return glm::quat(q[0], q[1], q[2], q[3]); }
void storeq(const glm::quat& q, float* o) { o[0] = q.w, o[1] = q.x, o[2] = q.y, o[3] = q.z; }
} // namespace

extern "C" {

void gs_mat3_mul_vec(const float* A, const float* v, float* out) {
    const glm::vec3 r = load<3, 3>(A) * glm::make_vec3(v);
    out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
void gs_mat3_mul_mat3(const float* A, const float* B, float* out) { store<3, 3>(load<3, 3>(A) * load<3, 3>(B), out); }
// (2x3 matrix: 3 columns, 2 rows) * (3x3) -> 2x3, then * transpose -> 2x2: the J * cov * J^T shape of the reference
void gs_mat3x2_chain(const float* J /* [2][3] */, const float* S /* [3][3] */, float* out /* [2][2] */) {
    const glm::mat<3, 2, float> j = load<3, 2>(J);
    store<2, 2>(j * load<3, 3>(S) * glm::transpose(j), out);
}
void gs_transpose3(const float* A, float* out) { store<3, 3>(glm::transpose(load<3, 3>(A)), out); }
void gs_outer3(const float* c, const float* r, float* out) { store<3, 3>(glm::outerProduct(glm::make_vec3(c), glm::make_vec3(r)), out); }
void gs_inverse2(const float* A, float* out) { store<2, 2>(glm::inverse(load<2, 2>(A)), out); }
void gs_mat3_arith(const float* A, const float* B, float s, float* sum, float* diff, float* neg, float* scaled, float* acc) {
    const glm::mat3 a = load<3, 3>(A), b = load<3, 3>(B);
    store<3, 3>(a + b, sum);
    store<3, 3>(a - b, diff);
    store<3, 3>(-a, neg);
    store<3, 3>(s * a, scaled);
    glm::mat3 c = a;
    c += b;
    store<3, 3>(c, acc);
}
// scalar-list constructor (column-major argument order), diagonal constructor, column constructor
void gs_mat3_ctors(const float* s9, float d, const float* c0, const float* c1, const float* c2, float* from_scalars, float* diag,
                   float* from_cols) {
    store<3, 3>(glm::mat3(s9[0], s9[1], s9[2], s9[3], s9[4], s9[5], s9[6], s9[7], s9[8]), from_scalars);
    store<3, 3>(glm::mat3(d), diag);
    store<3, 3>(glm::mat3(glm::make_vec3(c0), glm::make_vec3(c1), glm::make_vec3(c2)), from_cols);
}
void gs_vec3_ops(const float* a, const float* b, float* cross, float* dot, float* len, float* nrm) {
    const glm::vec3 x = glm::make_vec3(a), y = glm::make_vec3(b);
    const glm::vec3 c = glm::cross(x, y), n = glm::normalize(x);
    cross[0] = c.x, cross[1] = c.y, cross[2] = c.z;
    *dot = glm::dot(x, y);
    *len = glm::length(x);
    nrm[0] = n.x, nrm[1] = n.y, nrm[2] = n.z;
}
void gs_quat_rotate(const float* q, const float* v, float* by_operator, float* by_rotate) {
    const glm::vec3 a = loadq(q) * glm::make_vec3(v), b = glm::rotate(loadq(q), glm::make_vec3(v));
    by_operator[0] = a.x, by_operator[1] = a.y, by_operator[2] = a.z;
    by_rotate[0] = b.x, by_rotate[1] = b.y, by_rotate[2] = b.z;
}
void gs_mat3_cast(const float* q, float* out) { store<3, 3>(glm::mat3_cast(loadq(q)), out); }
void gs_quat_cast(const float* A, float* q) { storeq(glm::quat_cast(load<3, 3>(A)), q); }
void gs_quat_misc(const float* q, float* normalized, float* conj, float* inv, float* len) {
    storeq(glm::normalize(loadq(q)), normalized);
    storeq(glm::conjugate(loadq(q)), conj);
    storeq(glm::inverse(loadq(q)), inv);
    *len = glm::length(loadq(q));
}
void gs_slerp(const float* x, const float* y, float a, float* out) { storeq(glm::slerp(loadq(x), loadq(y), a), out); }
void gs_default_quat(float* out) { storeq(glm::quat(), out); }
float gs_mix(float x, float y, float a) { return glm::mix(x, y, a); }
}
