// Two-component vector stand-in for <glm/vec2.hpp>, used only when GLM itself is not on the include path (put
// include/compat AFTER the real GLM directory). Provides what STPNearestNeighbourInformation needs: x, y, constexpr
// construction, component-wise * and /.
#pragma once
namespace glm {
	template<typename T>
	struct tvec2 {
		T x, y;
		constexpr tvec2() : x(T(0)), y(T(0)) { }
		constexpr explicit tvec2(T v) : x(v), y(v) { }
		constexpr tvec2(T a, T b) : x(a), y(b) { }
	};
	template<typename T> constexpr tvec2<T> operator*(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x * b.x, a.y * b.y); }
	template<typename T> constexpr tvec2<T> operator*(const tvec2<T>& a, T k) { return tvec2<T>(a.x * k, a.y * k); }
	template<typename T> constexpr tvec2<T> operator/(const tvec2<T>& a, T k) { return tvec2<T>(a.x / k, a.y / k); }
	template<typename T> constexpr bool operator==(const tvec2<T>& a, const tvec2<T>& b) { return a.x == b.x && a.y == b.y; }
	using uvec2 = tvec2<unsigned int>;
	using vec2 = tvec2<float>;
	using ivec2 = tvec2<int>;
}
