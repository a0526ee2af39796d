// TEST INFRASTRUCTURE ONLY.
// Minimal glm-compatible shim so the UNMODIFIED reference gsplat backend (/root/reference/gsplat/*.cu)
// can be compiled in this image, where glm (a vcpkg dependency of the reference, vcpkg.json:8,
// version unpinned in-tree) is absent.  This is NOT glm: it provides only the ~25 symbols the
// reference uses, restating glm's published semantics (column-major mat<C,R>, m[col][row],
// qua{w,x,y,z}; rotate(q,v) = v + 2(w (u x v) + u x (u x v)); quat_cast = "biggest of four";
// slerp falls back to lerp when cos(theta) > 1 - epsilon).  Numerics follow the same operation order
// as upstream glm (detail/type_quat.inl, gtc/quaternion.inl) so results agree to rounding.
#pragma once
#include <cmath>
#include <cstddef>
#include <limits>

#if defined(__CUDACC__)
#define GLMS_FUNC __host__ __device__ inline
#else
#define GLMS_FUNC inline
#endif

namespace glm {

    typedef int length_t;

    template <length_t L, typename T>
    struct vec;

    template <typename T>
    struct vec<2, T> {
        T x, y;
        GLMS_FUNC vec() : x(0), y(0) {}
        GLMS_FUNC explicit vec(T s) : x(s), y(s) {}
        GLMS_FUNC vec(T a, T b) : x(a), y(b) {}
        template <typename A, typename B>
        GLMS_FUNC vec(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
        template <typename U>
        GLMS_FUNC vec(const vec<2, U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
        GLMS_FUNC T& operator[](length_t i) { return (&x)[i]; }
        GLMS_FUNC const T& operator[](length_t i) const { return (&x)[i]; }
    };

    template <typename T>
    struct vec<3, T> {
        T x, y, z;
        GLMS_FUNC vec() : x(0), y(0), z(0) {}
        GLMS_FUNC explicit vec(T s) : x(s), y(s), z(s) {}
        GLMS_FUNC vec(T a, T b, T c) : x(a), y(b), z(c) {}
        template <typename A, typename B, typename C>
        GLMS_FUNC vec(A a, B b, C c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
        template <typename U>
        GLMS_FUNC vec(const vec<3, U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
        GLMS_FUNC T& operator[](length_t i) { return (&x)[i]; }
        GLMS_FUNC const T& operator[](length_t i) const { return (&x)[i]; }
    };

    template <typename T>
    struct vec<4, T> {
        T x, y, z, w;
        GLMS_FUNC vec() : x(0), y(0), z(0), w(0) {}
        GLMS_FUNC explicit vec(T s) : x(s), y(s), z(s), w(s) {}
        GLMS_FUNC vec(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
        template <typename U>
        GLMS_FUNC vec(const vec<4, U>& v)
            : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)), w(static_cast<T>(v.w)) {}
        GLMS_FUNC T& operator[](length_t i) { return (&x)[i]; }
        GLMS_FUNC const T& operator[](length_t i) const { return (&x)[i]; }
    };

    // ---- component-wise vector arithmetic
#define GLMS_VEC_BINOP(OP)                                                                  \
    template <typename T>                                                                   \
    GLMS_FUNC vec<2, T> operator OP(const vec<2, T>& a, const vec<2, T>& b) {               \
        return vec<2, T>(a.x OP b.x, a.y OP b.y);                                           \
    }                                                                                       \
    template <typename T>                                                                   \
    GLMS_FUNC vec<3, T> operator OP(const vec<3, T>& a, const vec<3, T>& b) {               \
        return vec<3, T>(a.x OP b.x, a.y OP b.y, a.z OP b.z);                               \
    }                                                                                       \
    template <typename T>                                                                   \
    GLMS_FUNC vec<4, T> operator OP(const vec<4, T>& a, const vec<4, T>& b) {               \
        return vec<4, T>(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w);                   \
    }                                                                                       \
    template <typename T>                                                                   \
    GLMS_FUNC vec<2, T> operator OP(const vec<2, T>& a, T s) {                              \
        return vec<2, T>(a.x OP s, a.y OP s);                                               \
    }                                                                                       \
    template <typename T>                                                                   \
    GLMS_FUNC vec<3, T> operator OP(const vec<3, T>& a, T s) {                              \
        return vec<3, T>(a.x OP s, a.y OP s, a.z OP s);                                     \
    }                                                                                       \
    template <typename T>                                                                   \
    GLMS_FUNC vec<4, T> operator OP(const vec<4, T>& a, T s) {                              \
        return vec<4, T>(a.x OP s, a.y OP s, a.z OP s, a.w OP s);                           \
    }                                                                                       \
    template <typename T>                                                                   \
    GLMS_FUNC vec<2, T> operator OP(T s, const vec<2, T>& a) {                              \
        return vec<2, T>(s OP a.x, s OP a.y);                                               \
    }                                                                                       \
    template <typename T>                                                                   \
    GLMS_FUNC vec<3, T> operator OP(T s, const vec<3, T>& a) {                              \
        return vec<3, T>(s OP a.x, s OP a.y, s OP a.z);                                     \
    }                                                                                       \
    template <typename T>                                                                   \
    GLMS_FUNC vec<4, T> operator OP(T s, const vec<4, T>& a) {                              \
        return vec<4, T>(s OP a.x, s OP a.y, s OP a.z, s OP a.w);                           \
    }
    GLMS_VEC_BINOP(+)
    GLMS_VEC_BINOP(-)
    GLMS_VEC_BINOP(*)
    GLMS_VEC_BINOP(/)
#undef GLMS_VEC_BINOP

    template <typename T>
    GLMS_FUNC vec<2, T> operator-(const vec<2, T>& a) { return vec<2, T>(-a.x, -a.y); }
    template <typename T>
    GLMS_FUNC vec<3, T> operator-(const vec<3, T>& a) { return vec<3, T>(-a.x, -a.y, -a.z); }
    template <typename T>
    GLMS_FUNC vec<4, T> operator-(const vec<4, T>& a) { return vec<4, T>(-a.x, -a.y, -a.z, -a.w); }

#define GLMS_VEC_ASSIGN(OP)                                                        \
    template <length_t L, typename T>                                              \
    GLMS_FUNC vec<L, T>& operator OP##=(vec<L, T>& a, const vec<L, T>& b) {        \
        a = a OP b;                                                                \
        return a;                                                                  \
    }                                                                              \
    template <length_t L, typename T>                                              \
    GLMS_FUNC vec<L, T>& operator OP##=(vec<L, T>& a, T s) {                       \
        a = a OP s;                                                                \
        return a;                                                                  \
    }
    GLMS_VEC_ASSIGN(+)
    GLMS_VEC_ASSIGN(-)
    GLMS_VEC_ASSIGN(*)
    GLMS_VEC_ASSIGN(/)
#undef GLMS_VEC_ASSIGN

    template <typename T>
    GLMS_FUNC T dot(const vec<2, T>& a, const vec<2, T>& b) { return a.x * b.x + a.y * b.y; }
    template <typename T>
    GLMS_FUNC T dot(const vec<3, T>& a, const vec<3, T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    template <typename T>
    GLMS_FUNC T dot(const vec<4, T>& a, const vec<4, T>& b) {
        return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
    }
    template <typename T>
    GLMS_FUNC vec<3, T> cross(const vec<3, T>& x, const vec<3, T>& y) {
        return vec<3, T>(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
    }
    template <length_t L, typename T>
    GLMS_FUNC T length(const vec<L, T>& v) { return sqrt(dot(v, v)); }
    template <length_t L, typename T>
    GLMS_FUNC vec<L, T> normalize(const vec<L, T>& v) { return v * (static_cast<T>(1) / sqrt(dot(v, v))); }

    // ---- matrices: mat<C,R,T> has C columns of vec<R,T>; m[c][r]
    template <length_t C, length_t R, typename T>
    struct mat {
        typedef vec<R, T> col_type;
        col_type value[C];
        GLMS_FUNC mat() {
            for (length_t c = 0; c < C; ++c)
                value[c] = col_type();
        }
        // diagonal constructor
        GLMS_FUNC explicit mat(T s) {
            for (length_t c = 0; c < C; ++c) {
                value[c] = col_type();
                if (c < R)
                    value[c][c] = s;
            }
        }
        // column-major scalar lists
        GLMS_FUNC mat(T a0, T a1, T a2, T a3) {
            static_assert(C * R == 4, "scalar ctor arity");
            T a[4] = {a0, a1, a2, a3};
            fill(a);
        }
        GLMS_FUNC mat(T a0, T a1, T a2, T a3, T a4, T a5) {
            static_assert(C * R == 6, "scalar ctor arity");
            T a[6] = {a0, a1, a2, a3, a4, a5};
            fill(a);
        }
        GLMS_FUNC mat(T a0, T a1, T a2, T a3, T a4, T a5, T a6, T a7, T a8) {
            static_assert(C * R == 9, "scalar ctor arity");
            T a[9] = {a0, a1, a2, a3, a4, a5, a6, a7, a8};
            fill(a);
        }
        GLMS_FUNC mat(const col_type& c0, const col_type& c1) {
            static_assert(C == 2, "column ctor arity");
            value[0] = c0;
            value[1] = c1;
        }
        GLMS_FUNC mat(const col_type& c0, const col_type& c1, const col_type& c2) {
            static_assert(C == 3, "column ctor arity");
            value[0] = c0;
            value[1] = c1;
            value[2] = c2;
        }
        GLMS_FUNC col_type& operator[](length_t i) { return value[i]; }
        GLMS_FUNC const col_type& operator[](length_t i) const { return value[i]; }
        // mat * column vector (vec<C>) -> vec<R>; hidden friend so the vector operand converts implicitly
        friend GLMS_FUNC vec<R, T> operator*(const mat& m, const vec<C, T>& v) {
            vec<R, T> r = m.value[0] * v[0];
            for (length_t c = 1; c < C; ++c)
                r = r + m.value[c] * v[c];
            return r;
        }

    private:
        GLMS_FUNC void fill(const T* a) {
            for (length_t c = 0; c < C; ++c)
                for (length_t r = 0; r < R; ++r)
                    value[c][r] = a[c * R + r];
        }
    };

    template <length_t C, length_t R, typename T>
    GLMS_FUNC mat<C, R, T> operator+(const mat<C, R, T>& a, const mat<C, R, T>& b) {
        mat<C, R, T> m;
        for (length_t c = 0; c < C; ++c)
            m[c] = a[c] + b[c];
        return m;
    }
    template <length_t C, length_t R, typename T>
    GLMS_FUNC mat<C, R, T> operator-(const mat<C, R, T>& a, const mat<C, R, T>& b) {
        mat<C, R, T> m;
        for (length_t c = 0; c < C; ++c)
            m[c] = a[c] - b[c];
        return m;
    }
    template <length_t C, length_t R, typename T>
    GLMS_FUNC mat<C, R, T> operator-(const mat<C, R, T>& a) {
        mat<C, R, T> m;
        for (length_t c = 0; c < C; ++c)
            m[c] = -a[c];
        return m;
    }
    template <length_t C, length_t R, typename T>
    GLMS_FUNC mat<C, R, T> operator*(const mat<C, R, T>& a, T s) {
        mat<C, R, T> m;
        for (length_t c = 0; c < C; ++c)
            m[c] = a[c] * s;
        return m;
    }
    template <length_t C, length_t R, typename T>
    GLMS_FUNC mat<C, R, T> operator*(T s, const mat<C, R, T>& a) {
        mat<C, R, T> m;
        for (length_t c = 0; c < C; ++c)
            m[c] = s * a[c];
        return m;
    }
    template <length_t C, length_t R, typename T>
    GLMS_FUNC mat<C, R, T>& operator+=(mat<C, R, T>& a, const mat<C, R, T>& b) {
        for (length_t c = 0; c < C; ++c)
            a[c] = a[c] + b[c];
        return a;
    }
    // mat<K,R> * mat<C,K> -> mat<C,R>
    template <length_t K, length_t R, length_t C, typename T>
    GLMS_FUNC mat<C, R, T> operator*(const mat<K, R, T>& a, const mat<C, K, T>& b) {
        mat<C, R, T> m;
        for (length_t c = 0; c < C; ++c)
            m[c] = a * b[c];
        return m;
    }
    template <length_t C, length_t R, typename T>
    GLMS_FUNC mat<R, C, T> transpose(const mat<C, R, T>& a) {
        mat<R, C, T> m;
        for (length_t c = 0; c < C; ++c)
            for (length_t r = 0; r < R; ++r)
                m[r][c] = a[c][r];
        return m;
    }
    // outerProduct(c, r) = c * r^T  -> mat<len(r), len(c)>
    template <length_t DA, length_t DB, typename T>
    GLMS_FUNC mat<DB, DA, T> outerProduct(const vec<DA, T>& c, const vec<DB, T>& r) {
        mat<DB, DA, T> m;
        for (length_t i = 0; i < DB; ++i)
            m[i] = c * r[i];
        return m;
    }
    template <typename T>
    GLMS_FUNC mat<2, 2, T> inverse(const mat<2, 2, T>& m) {
        T OneOverDeterminant = static_cast<T>(1) / (m[0][0] * m[1][1] - m[1][0] * m[0][1]);
        return mat<2, 2, T>(
            +m[1][1] * OneOverDeterminant, -m[0][1] * OneOverDeterminant, -m[1][0] * OneOverDeterminant,
            +m[0][0] * OneOverDeterminant);
    }

    // ---- quaternion (w, x, y, z constructor order; storage order irrelevant to callers)
    template <typename T>
    struct qua {
        T x, y, z, w;
        GLMS_FUNC qua() : x(0), y(0), z(0), w(1) {}
        GLMS_FUNC qua(T w_, T x_, T y_, T z_) : x(x_), y(y_), z(z_), w(w_) {}
    };
    template <typename T>
    GLMS_FUNC T dot(const qua<T>& a, const qua<T>& b) {
        return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
    }
    template <typename T>
    GLMS_FUNC T length(const qua<T>& q) { return sqrt(dot(q, q)); }
    template <typename T>
    GLMS_FUNC qua<T> normalize(const qua<T>& q) {
        T len = length(q);
        if (len <= static_cast<T>(0))
            return qua<T>(static_cast<T>(1), static_cast<T>(0), static_cast<T>(0), static_cast<T>(0));
        T oneOverLen = static_cast<T>(1) / len;
        return qua<T>(q.w * oneOverLen, q.x * oneOverLen, q.y * oneOverLen, q.z * oneOverLen);
    }
    template <typename T>
    GLMS_FUNC qua<T> conjugate(const qua<T>& q) { return qua<T>(q.w, -q.x, -q.y, -q.z); }
    template <typename T>
    GLMS_FUNC qua<T> inverse(const qua<T>& q) {
        T d = dot(q, q);
        qua<T> c = conjugate(q);
        return qua<T>(c.w / d, c.x / d, c.y / d, c.z / d);
    }
    template <typename T>
    GLMS_FUNC qua<T> operator-(const qua<T>& q) { return qua<T>(-q.w, -q.x, -q.y, -q.z); }
    template <typename T>
    GLMS_FUNC qua<T> operator+(const qua<T>& a, const qua<T>& b) {
        return qua<T>(a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z);
    }
    template <typename T>
    GLMS_FUNC qua<T> operator*(const qua<T>& a, T s) { return qua<T>(a.w * s, a.x * s, a.y * s, a.z * s); }
    template <typename T>
    GLMS_FUNC qua<T> operator*(T s, const qua<T>& a) { return a * s; }
    template <typename T>
    GLMS_FUNC qua<T> operator/(const qua<T>& a, T s) { return qua<T>(a.w / s, a.x / s, a.y / s, a.z / s); }
    // q * v  (rotate a vector)
    template <typename T>
    GLMS_FUNC vec<3, T> operator*(const qua<T>& q, const vec<3, T>& v) {
        const vec<3, T> QuatVector(q.x, q.y, q.z);
        const vec<3, T> uv(cross(QuatVector, v));
        const vec<3, T> uuv(cross(QuatVector, uv));
        return v + ((uv * q.w) + uuv) * static_cast<T>(2);
    }
    template <typename T>
    GLMS_FUNC vec<3, T> rotate(const qua<T>& q, const vec<3, T>& v) { return q * v; }

    template <typename T>
    GLMS_FUNC mat<3, 3, T> mat3_cast(const qua<T>& q) {
        mat<3, 3, T> Result(static_cast<T>(1));
        T qxx(q.x * q.x);
        T qyy(q.y * q.y);
        T qzz(q.z * q.z);
        T qxz(q.x * q.z);
        T qxy(q.x * q.y);
        T qyz(q.y * q.z);
        T qwx(q.w * q.x);
        T qwy(q.w * q.y);
        T qwz(q.w * q.z);
        Result[0][0] = T(1) - T(2) * (qyy + qzz);
        Result[0][1] = T(2) * (qxy + qwz);
        Result[0][2] = T(2) * (qxz - qwy);
        Result[1][0] = T(2) * (qxy - qwz);
        Result[1][1] = T(1) - T(2) * (qxx + qzz);
        Result[1][2] = T(2) * (qyz + qwx);
        Result[2][0] = T(2) * (qxz + qwy);
        Result[2][1] = T(2) * (qyz - qwx);
        Result[2][2] = T(1) - T(2) * (qxx + qyy);
        return Result;
    }

    template <typename T>
    GLMS_FUNC qua<T> quat_cast(const mat<3, 3, T>& m) {
        T fourXSquaredMinus1 = m[0][0] - m[1][1] - m[2][2];
        T fourYSquaredMinus1 = m[1][1] - m[0][0] - m[2][2];
        T fourZSquaredMinus1 = m[2][2] - m[0][0] - m[1][1];
        T fourWSquaredMinus1 = m[0][0] + m[1][1] + m[2][2];
        int biggestIndex = 0;
        T fourBiggestSquaredMinus1 = fourWSquaredMinus1;
        if (fourXSquaredMinus1 > fourBiggestSquaredMinus1) {
            fourBiggestSquaredMinus1 = fourXSquaredMinus1;
            biggestIndex = 1;
        }
        if (fourYSquaredMinus1 > fourBiggestSquaredMinus1) {
            fourBiggestSquaredMinus1 = fourYSquaredMinus1;
            biggestIndex = 2;
        }
        if (fourZSquaredMinus1 > fourBiggestSquaredMinus1) {
            fourBiggestSquaredMinus1 = fourZSquaredMinus1;
            biggestIndex = 3;
        }
        T biggestVal = sqrt(fourBiggestSquaredMinus1 + static_cast<T>(1)) * static_cast<T>(0.5);
        T mult = static_cast<T>(0.25) / biggestVal;
        switch (biggestIndex) {
        case 0:
            return qua<T>(biggestVal, (m[1][2] - m[2][1]) * mult, (m[2][0] - m[0][2]) * mult, (m[0][1] - m[1][0]) * mult);
        case 1:
            return qua<T>((m[1][2] - m[2][1]) * mult, biggestVal, (m[0][1] + m[1][0]) * mult, (m[2][0] + m[0][2]) * mult);
        case 2:
            return qua<T>((m[2][0] - m[0][2]) * mult, (m[0][1] + m[1][0]) * mult, biggestVal, (m[1][2] + m[2][1]) * mult);
        default:
            return qua<T>((m[0][1] - m[1][0]) * mult, (m[2][0] + m[0][2]) * mult, (m[1][2] + m[2][1]) * mult, biggestVal);
        }
    }

    template <typename T>
    GLMS_FUNC T mix(T x, T y, T a) { return x * (static_cast<T>(1) - a) + y * a; }

    template <typename T>
    GLMS_FUNC qua<T> slerp(const qua<T>& x, const qua<T>& y, T a) {
        qua<T> z = y;
        T cosTheta = dot(x, y);
        // take the short way round
        if (cosTheta < static_cast<T>(0)) {
            z = -y;
            cosTheta = -cosTheta;
        }
        if (cosTheta > static_cast<T>(1) - std::numeric_limits<T>::epsilon()) {
            return qua<T>(mix(x.w, z.w, a), mix(x.x, z.x, a), mix(x.y, z.y, a), mix(x.z, z.z, a));
        } else {
            T angle = acos(cosTheta);
            return (sin((static_cast<T>(1) - a) * angle) * x + sin(a * angle) * z) / sin(angle);
        }
    }

    // ---- type_ptr
    template <typename T>
    GLMS_FUNC vec<2, T> make_vec2(const T* p) { return vec<2, T>(p[0], p[1]); }
    template <typename T>
    GLMS_FUNC vec<3, T> make_vec3(const T* p) { return vec<3, T>(p[0], p[1], p[2]); }
    template <typename T>
    GLMS_FUNC vec<4, T> make_vec4(const T* p) { return vec<4, T>(p[0], p[1], p[2], p[3]); }

    // ---- aliases used by the reference
    typedef vec<2, float> vec2;
    typedef vec<3, float> vec3;
    typedef vec<4, float> vec4;
    typedef vec<2, float> fvec2;
    typedef vec<3, float> fvec3;
    typedef vec<4, float> fvec4;
    typedef mat<2, 2, float> mat2;
    typedef mat<3, 3, float> mat3;
    typedef mat<4, 4, float> mat4;
    typedef mat<2, 2, float> fmat2;
    typedef mat<3, 3, float> fmat3;
    typedef mat<4, 4, float> fmat4;
    typedef qua<float> quat;
    typedef qua<float> fquat;

} // namespace glm
