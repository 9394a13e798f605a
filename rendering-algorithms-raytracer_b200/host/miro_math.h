// miro_math.h — minimal host-side vector / matrix types of the product's C++ host layer.
// Mirrors the parts of the reference's Vector3.h / Matrix4x4.h API that scene set-up uses
// (row-major 4x4, m_ij = row i column j, src/Matrix4x4.h:20-26; translate/scale/rotate act on the
// left as in src/Matrix4x4.h:751-856).  Plain scalar FP32 — no SSE approximations.
#pragma once
#include <math.h>
#include <string.h>

namespace miro {

constexpr float PI = 3.1415926f;   // src/Miro.h:57

struct Vector3 {
    float x, y, z;
    Vector3() : x(0), y(0), z(0) {}
    explicit Vector3(float s) : x(s), y(s), z(s) {}
    Vector3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    Vector3 operator+(const Vector3& o) const { return Vector3(x + o.x, y + o.y, z + o.z); }
    Vector3 operator-(const Vector3& o) const { return Vector3(x - o.x, y - o.y, z - o.z); }
    Vector3 operator-() const { return Vector3(-x, -y, -z); }
    Vector3 operator*(float s) const { return Vector3(x * s, y * s, z * s); }
    float length2() const { return x * x + y * y + z * z; }
    float length() const { return sqrtf(length2()); }
    Vector3 normalized() const { float l = 1.0f / sqrtf(length2()); return Vector3(x * l, y * l, z * l); }
    void normalize() { *this = normalized(); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vector3 operator*(float s, const Vector3& a) { return a * s; }
inline float dot(const Vector3& a, const Vector3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vector3 cross(const Vector3& a, const Vector3& b) { return Vector3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

struct Matrix4x4 {
    float m[16];   // row-major: m[4*r + c]
    Matrix4x4() { setIdentity(); }
    void setIdentity() { memset(m, 0, sizeof(m)); m[0] = m[5] = m[10] = m[15] = 1.f; }
    static Matrix4x4 fromRowMajor(const float* v) { Matrix4x4 r; memcpy(r.m, v, sizeof(r.m)); return r; }
    float& at(int r, int c) { return m[4 * r + c]; }
    float at(int r, int c) const { return m[4 * r + c]; }
    Matrix4x4 operator*(const Matrix4x4& b) const {
        Matrix4x4 r;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += at(i, k) * b.at(k, j);
            r.at(i, j) = s;
        }
        return r;
    }
    // src/Matrix4x4.h:751-778: translate adds to column 4; scale multiplies the diagonal-by-rows
    void translate(float x, float y, float z) { at(0, 3) += x; at(1, 3) += y; at(2, 3) += z; }
    void scale(float x, float y, float z);
    void rotate(float angleDeg, float x, float y, float z);
    Matrix4x4 transposed() const { Matrix4x4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.at(i, j) = at(j, i); return r; }
    bool isAffine() const { return m[12] == 0.f && m[13] == 0.f && m[14] == 0.f && m[15] == 1.f; }
    // general inverse (double precision cofactors); returns false when singular
    bool inverted(Matrix4x4& out) const;
    // the inverse with the ROUNDING of the reference's Matrix4x4::invert (src/Matrix4x4.h:354-411): single-precision cofactor
    // expansion over 2x2 and 3x3 sub-determinants, evaluated left to right.  A ProxyObject's object-space ray is
    // M^-1 applied to the world-space ray (src/ProxyObject.cpp:78-79); for hit distances to agree with the reference to the
    // last bits the matrix entries must be the reference's, not a better inverse.  (Compiled with -ffp-contract=off.)
    bool invertedAsReference(Matrix4x4& out) const;
    Vector3 transformPoint(const Vector3& p) const {   // multiplyAndDivideByW, src/Matrix4x4.h:714-749
        float w = at(3, 0) * p.x + at(3, 1) * p.y + at(3, 2) * p.z + at(3, 3);
        float iw = 1.0f / w;
        return Vector3((at(0, 0) * p.x + at(0, 1) * p.y + at(0, 2) * p.z + at(0, 3)) * iw,
                       (at(1, 0) * p.x + at(1, 1) * p.y + at(1, 2) * p.z + at(1, 3)) * iw,
                       (at(2, 0) * p.x + at(2, 1) * p.y + at(2, 2) * p.z + at(2, 3)) * iw);
    }
    Vector3 transformVector(const Vector3& v) const {  // operator*(Matrix4x4, Vector3): ignores row 4 and column 4
        return Vector3(at(0, 0) * v.x + at(0, 1) * v.y + at(0, 2) * v.z,
                       at(1, 0) * v.x + at(1, 1) * v.y + at(1, 2) * v.z,
                       at(2, 0) * v.x + at(2, 1) * v.y + at(2, 2) * v.z);
    }
};

// src/Matrix4x4.h:757-762: scale touches the three diagonal entries only (quirk kept: after a
// rotate() this is not a general scaling)
inline void Matrix4x4::scale(float x, float y, float z) { at(0, 0) *= x; at(1, 1) *= y; at(2, 2) *= z; }

// src/Matrix4x4.h:831-856: rotate() OVERWRITES the matrix with an axis-angle rotation ("set" fills by columns)
inline void Matrix4x4::rotate(float angleDeg, float x, float y, float z) {
    float rad = angleDeg * (PI / 180.);
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float c = cos(rad), cinv = 1 - c, s = sin(rad);
    float xy = x * y, xz = x * z, yz = y * z, xs = x * s, ys = y * s, zs = z * s;
    float xzc = xz * cinv, xyc = xy * cinv, yzc = yz * cinv;
    const float cols[16] = {x2 + c * (1 - x2), xy * cinv + zs, xzc - ys, 0,
                            xyc - zs, y2 + c * (1 - y2), yzc + xs, 0,
                            xzc + ys, yzc - xs, z2 + c * (1 - z2), 0,
                            0, 0, 0, 1};
    for (int col = 0; col < 4; ++col) for (int row = 0; row < 4; ++row) at(row, col) = cols[4 * col + row];
}

inline bool Matrix4x4::inverted(Matrix4x4& out) const {
    double a[16], inv[16];
    for (int i = 0; i < 16; ++i) a[i] = m[i];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    if (det == 0.0) return false;
    det = 1.0 / det;
    for (int i = 0; i < 16; ++i) out.m[i] = (float)(inv[i] * det);
    return true;
}

inline bool Matrix4x4::invertedAsReference(Matrix4x4& out) const {
    // minor2(r, s, c, e): rows r < s, columns c < e (0-based) of *this
    auto minor2 = [&](int r, int s, int c, int e) { return at(r, c) * at(s, e) - at(r, e) * at(s, c); };
    // minor3(i, j): the sub-determinant without row i and column j, expanded along its first remaining row over the 2x2
    // minors of its last two rows: first term minus second plus third, in that order
    auto minor3 = [&](int i, int j) {
        int rows[3], cols[3];
        for (int k = 0, n = 0; k < 4; ++k) if (k != i) rows[n++] = k;
        for (int k = 0, n = 0; k < 4; ++k) if (k != j) cols[n++] = k;
        const int r = rows[0], p = rows[1], q = rows[2];
        return at(r, cols[0]) * minor2(p, q, cols[1], cols[2]) - at(r, cols[1]) * minor2(p, q, cols[0], cols[2]) + at(r, cols[2]) * minor2(p, q, cols[0], cols[1]);
    };
    float sd[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) sd[i][j] = minor3(i, j);
    const float det = at(0, 0) * sd[0][0] - at(0, 1) * sd[0][1] + at(0, 2) * sd[0][2] - at(0, 3) * sd[0][3];
    if (det == 0.0f) return false;
    const float detInv = (float)(1.0 / (double)det);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out.at(i, j) = (((i + j) & 1) ? -sd[j][i] : sd[j][i]) * detInv;
    return true;
}

// What the reference multiplies a transformed point by (Matrix4x4::multiplyAndDivideByW, src/Matrix4x4.h:728-733):
// recipps(w) = rcpps(w) refined by one Newton step (src/SSE.h:81-86) — evaluated on THIS host's SSE unit, because the
// table behind rcpps differs between CPU vendors and the reference running on this host would use this host's.
float referenceRecip(float w);

}  // namespace miro
