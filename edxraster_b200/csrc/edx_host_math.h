// edx_host_math.h — host-side matrix helpers of Renderer::SetTransform (Core/Renderer.cpp:85-92).
//
// The reference gets these from EDXUtil's Matrix class, which is not available, so the
// arithmetic is defined here (DESIGN.md "EDXUtil definitions" 2 and 4): plain fp32, products summed left
// to right, no FMA contraction (the library is built with -ffp-contract=off). Header-only so the C ABI
// and the C++ host API share one definition.
#pragma once
#include <cmath>
#include <cstring>

namespace edx_host {

struct Mat4 {
    float m[16];                                   // row-major, column-vector convention
    float at(int r, int c) const { return m[4 * r + c]; }
    float& at(int r, int c) { return m[4 * r + c]; }
};

inline Mat4 identity()
{
    Mat4 r;
    for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    return r;
}

// Matrix operator* : r(i,j) = ((a(i,0) b(0,j) + a(i,1) b(1,j)) + a(i,2) b(2,j)) + a(i,3) b(3,j)
inline Mat4 multiply(const Mat4& a, const Mat4& b)
{
    Mat4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float acc = a.at(i, 0) * b.at(0, j) + a.at(i, 1) * b.at(1, j);
            acc = acc + a.at(i, 2) * b.at(2, j);
            acc = acc + a.at(i, 3) * b.at(3, j);
            r.at(i, j) = acc;
        }
    return r;
}

// Matrix::Inverse by the adjugate, built from the twelve 2x2 minors of the top and bottom row pairs.
inline Mat4 inverse(const Mat4& a)
{
    const float* m = a.m;
    auto M = [&](int r, int c) { return m[4 * r + c]; };
    const float s0 = M(0, 0) * M(1, 1) - M(1, 0) * M(0, 1);
    const float s1 = M(0, 0) * M(1, 2) - M(1, 0) * M(0, 2);
    const float s2 = M(0, 0) * M(1, 3) - M(1, 0) * M(0, 3);
    const float s3 = M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2);
    const float s4 = M(0, 1) * M(1, 3) - M(1, 1) * M(0, 3);
    const float s5 = M(0, 2) * M(1, 3) - M(1, 2) * M(0, 3);
    const float c5 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    const float c4 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    const float c3 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    const float c2 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    const float c1 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
    const float c0 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    const float det = ((((s0 * c5 - s1 * c4) + s2 * c3) + s3 * c2) - s4 * c1) + s5 * c0;
    const float id = 1.0f / det;
    Mat4 r;
    r.at(0, 0) = ((M(1, 1) * c5 - M(1, 2) * c4) + M(1, 3) * c3) * id;
    r.at(0, 1) = ((-M(0, 1) * c5 + M(0, 2) * c4) - M(0, 3) * c3) * id;
    r.at(0, 2) = ((M(3, 1) * s5 - M(3, 2) * s4) + M(3, 3) * s3) * id;
    r.at(0, 3) = ((-M(2, 1) * s5 + M(2, 2) * s4) - M(2, 3) * s3) * id;
    r.at(1, 0) = ((-M(1, 0) * c5 + M(1, 2) * c2) - M(1, 3) * c1) * id;
    r.at(1, 1) = ((M(0, 0) * c5 - M(0, 2) * c2) + M(0, 3) * c1) * id;
    r.at(1, 2) = ((-M(3, 0) * s5 + M(3, 2) * s2) - M(3, 3) * s1) * id;
    r.at(1, 3) = ((M(2, 0) * s5 - M(2, 2) * s2) + M(2, 3) * s1) * id;
    r.at(2, 0) = ((M(1, 0) * c4 - M(1, 1) * c2) + M(1, 3) * c0) * id;
    r.at(2, 1) = ((-M(0, 0) * c4 + M(0, 1) * c2) - M(0, 3) * c0) * id;
    r.at(2, 2) = ((M(3, 0) * s4 - M(3, 1) * s2) + M(3, 3) * s0) * id;
    r.at(2, 3) = ((-M(2, 0) * s4 + M(2, 1) * s2) - M(2, 3) * s0) * id;
    r.at(3, 0) = ((-M(1, 0) * c3 + M(1, 1) * c1) - M(1, 2) * c0) * id;
    r.at(3, 1) = ((M(0, 0) * c3 - M(0, 1) * c1) + M(0, 2) * c0) * id;
    r.at(3, 2) = ((-M(3, 0) * s3 + M(3, 1) * s1) - M(3, 2) * s0) * id;
    r.at(3, 3) = ((M(2, 0) * s3 - M(2, 1) * s1) + M(2, 2) * s0) * id;
    return r;
}

// Matrix::TransformPoint(Vector3, M): w_in = 1, divide by w' only when w' != 1 (Renderer.cpp:289)
inline void transform_point3(const Mat4& M, const float in[3], float out[3])
{
    float x = ((M.at(0, 0) * in[0] + M.at(0, 1) * in[1]) + M.at(0, 2) * in[2]) + M.at(0, 3);
    float y = ((M.at(1, 0) * in[0] + M.at(1, 1) * in[1]) + M.at(1, 2) * in[2]) + M.at(1, 3);
    float z = ((M.at(2, 0) * in[0] + M.at(2, 1) * in[1]) + M.at(2, 2) * in[2]) + M.at(2, 3);
    float w = ((M.at(3, 0) * in[0] + M.at(3, 1) * in[1]) + M.at(3, 2) * in[2]) + M.at(3, 3);
    if (w != 1.0f) { x = x / w; y = y / w; z = z / w; }
    out[0] = x; out[1] = y; out[2] = z;
}

// Math::Normalize(Vector3) (Shader.h:258): v / |v| by true division
inline void normalize3(const float in[3], float out[3])
{
    float len = std::sqrt((in[0] * in[0] + in[1] * in[1]) + in[2] * in[2]);
    out[0] = in[0] / len; out[1] = in[1] / len; out[2] = in[2] / len;
}

} // namespace edx_host
