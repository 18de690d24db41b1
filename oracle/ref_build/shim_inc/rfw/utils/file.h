// stand-in for the parts of rfw/utils/file.h and rfw/utils/logger.h Camera.cpp names (never executed by the pin)
#pragma once
#include <cstdio>
#include <string_view>
#include <vector>
#define WARNING(...) ((void)0)
namespace rfw
{
namespace utils
{
namespace file
{
inline bool exists(std::string_view) { return false; }
inline std::vector<char> read_binary(std::string_view) { return {}; }
inline void write(std::string_view, const std::vector<char> &) {}
} // namespace file
} // namespace utils
} // namespace rfw
