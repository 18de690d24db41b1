// stand-in: Camera.cpp's (de)serialisation members only have to compile; the pin exercises get_view()
#pragma once
#include <memory>
#include <vector>
namespace rfw
{
namespace utils
{
template <typename T, int N> class serializable
{
  public:
	serializable() = default;
	explicit serializable(const T &) {}
	std::vector<char> serialize() const { return {}; }
	static serializable deserialize(const std::vector<char> &) { return serializable(); }
	std::shared_ptr<T> get_data() const { return std::make_shared<T>(); }
};
} // namespace utils
} // namespace rfw
