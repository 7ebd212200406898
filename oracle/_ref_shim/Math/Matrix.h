// EDXUtil stand-in (oracle/_ref_shim): Matrix, 4x4 fp32, m[row][col], column-vector convention
// (Renderer.cpp:90 `ModelViewProjMatrix = mProj * mModelView`). DESIGN.md shims 1, 2, 4.
#pragma once
#include "Vector.h"
namespace EDX
{
	class Matrix
	{
	public:
		float m[4][4];

		Matrix()
		{
			for (int i = 0; i < 4; i++)
				for (int j = 0; j < 4; j++)
					m[i][j] = i == j ? 1.0f : 0.0f;
		}
		explicit Matrix(const float* rowMajor16) { memcpy(m, rowMajor16, sizeof(m)); }

		// shim 4: row times column, summed left to right
		Matrix operator*(const Matrix& b) const
		{
			Matrix r;
			for (int i = 0; i < 4; i++)
				for (int j = 0; j < 4; j++)
					r.m[i][j] = ((m[i][0] * b.m[0][j] + m[i][1] * b.m[1][j]) + m[i][2] * b.m[2][j]) + m[i][3] * b.m[3][j];
			return r;
		}

		// shim 1: Matrix::TransformPoint(Vector4, M) (Shader.h:45): full 4x4, row . vector, left to right
		static Vector4 TransformPoint(const Vector4& p, const Matrix& M)
		{
			return Vector4(
				((M.m[0][0] * p.x + M.m[0][1] * p.y) + M.m[0][2] * p.z) + M.m[0][3] * p.w,
				((M.m[1][0] * p.x + M.m[1][1] * p.y) + M.m[1][2] * p.z) + M.m[1][3] * p.w,
				((M.m[2][0] * p.x + M.m[2][1] * p.y) + M.m[2][2] * p.z) + M.m[2][3] * p.w,
				((M.m[3][0] * p.x + M.m[3][1] * p.y) + M.m[3][2] * p.z) + M.m[3][3] * p.w);
		}

		// shim 2: Matrix::TransformPoint(Vector3, M) (RasterTriangle.h:30-32, Renderer.cpp:289): w = 1 implied, divide by
		// w' only when w' != 1
		static Vector3 TransformPoint(const Vector3& p, const Matrix& M)
		{
			float x = ((M.m[0][0] * p.x + M.m[0][1] * p.y) + M.m[0][2] * p.z) + M.m[0][3];
			float y = ((M.m[1][0] * p.x + M.m[1][1] * p.y) + M.m[1][2] * p.z) + M.m[1][3];
			float z = ((M.m[2][0] * p.x + M.m[2][1] * p.y) + M.m[2][2] * p.z) + M.m[2][3];
			float w = ((M.m[3][0] * p.x + M.m[3][1] * p.y) + M.m[3][2] * p.z) + M.m[3][3];
			if (w != 1.0f) { x = x / w; y = y / w; z = z / w; }
			return Vector3(x, y, z);
		}

		// shim 4: adjugate from 2x2 sub-determinants (host only: feeds the eye position, Renderer.cpp:289)
		static Matrix Inverse(const Matrix& a)
		{
			const float (*m)[4] = a.m;
			const float s0 = m[0][0] * m[1][1] - m[1][0] * m[0][1];
			const float s1 = m[0][0] * m[1][2] - m[1][0] * m[0][2];
			const float s2 = m[0][0] * m[1][3] - m[1][0] * m[0][3];
			const float s3 = m[0][1] * m[1][2] - m[1][1] * m[0][2];
			const float s4 = m[0][1] * m[1][3] - m[1][1] * m[0][3];
			const float s5 = m[0][2] * m[1][3] - m[1][2] * m[0][3];
			const float c5 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
			const float c4 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
			const float c3 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
			const float c2 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
			const float c1 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
			const float c0 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
			const float det = ((((s0 * c5 - s1 * c4) + s2 * c3) + s3 * c2) - s4 * c1) + s5 * c0;
			const float id = 1.0f / det;
			Matrix r;
			r.m[0][0] = ((m[1][1] * c5 - m[1][2] * c4) + m[1][3] * c3) * id;
			r.m[0][1] = ((-m[0][1] * c5 + m[0][2] * c4) - m[0][3] * c3) * id;
			r.m[0][2] = ((m[3][1] * s5 - m[3][2] * s4) + m[3][3] * s3) * id;
			r.m[0][3] = ((-m[2][1] * s5 + m[2][2] * s4) - m[2][3] * s3) * id;
			r.m[1][0] = ((-m[1][0] * c5 + m[1][2] * c2) - m[1][3] * c1) * id;
			r.m[1][1] = ((m[0][0] * c5 - m[0][2] * c2) + m[0][3] * c1) * id;
			r.m[1][2] = ((-m[3][0] * s5 + m[3][2] * s2) - m[3][3] * s1) * id;
			r.m[1][3] = ((m[2][0] * s5 - m[2][2] * s2) + m[2][3] * s1) * id;
			r.m[2][0] = ((m[1][0] * c4 - m[1][1] * c2) + m[1][3] * c0) * id;
			r.m[2][1] = ((-m[0][0] * c4 + m[0][1] * c2) - m[0][3] * c0) * id;
			r.m[2][2] = ((m[3][0] * s4 - m[3][1] * s2) + m[3][3] * s0) * id;
			r.m[2][3] = ((-m[2][0] * s4 + m[2][1] * s2) - m[2][3] * s0) * id;
			r.m[3][0] = ((-m[1][0] * c3 + m[1][1] * c1) - m[1][2] * c0) * id;
			r.m[3][1] = ((m[0][0] * c3 - m[0][1] * c1) + m[0][2] * c0) * id;
			r.m[3][2] = ((-m[3][0] * s3 + m[3][1] * s1) - m[3][2] * s0) * id;
			r.m[3][3] = ((m[2][0] * s3 - m[2][1] * s1) + m[2][2] * s0) * id;
			return r;
		}
	};
}
