#pragma once
#include <rfw/context/device_structs.h>
