// Stand-in used ONLY by the local adapter check build (host/Makefile): the reference's context.h includes
// <rfw/utils/window.h> (GLFW + GLEW) just for the `window` type and GLuint.  A maintainer building inside the
// reference tree uses the real header.
#pragma once
typedef unsigned int GLuint;
namespace rfw
{
namespace utils
{
class window;
}
} // namespace rfw
