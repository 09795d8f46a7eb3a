// Minimal stand-in for <glm/vec2.hpp> (GLM is not installed in this image). Own code, test infrastructure only:
// provides exactly the two-component vector operations the reference filter and biome factory touch
// (/root/reference/SuperTerrain+/SuperAlgorithm+/Host/Private/STPSingleHistogramFilter.cpp:876-878).
#pragma once
namespace glm {
template <typename T> struct tvec2 {
    T x, y;
    constexpr tvec2() : x(T(0)), y(T(0)) {}
    constexpr explicit tvec2(T v) : x(v), y(v) {}
    constexpr tvec2(T a, T b) : x(a), y(b) {}
    tvec2& operator*=(T k) { x *= k; y *= k; return *this; }
};
template <typename T> constexpr tvec2<T> operator*(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x * b.x, a.y * b.y); }
template <typename T> constexpr tvec2<T> operator*(const tvec2<T>& a, T k) { return tvec2<T>(a.x * k, a.y * k); }
template <typename T> constexpr tvec2<T> operator/(const tvec2<T>& a, T k) { return tvec2<T>(a.x / k, a.y / k); }
using uvec2 = tvec2<unsigned int>;
using vec2 = tvec2<float>;
using ivec2 = tvec2<int>;   // STPBiomeFactory::operator()(STPSample_t*, glm::ivec2)
}
