// Stand-in for the local adapter check build: camera.h only names this template in two declarations.
#pragma once
namespace rfw
{
namespace utils
{
template <typename T, int N> class serializable;
}
} // namespace rfw
