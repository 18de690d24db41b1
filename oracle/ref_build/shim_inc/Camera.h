// Camera.cpp says #include "Camera.h"; the file in the reference tree is camera.h (case-insensitive file systems)
#pragma once
#include <rfw/context/camera.h>
