// the reference includes "../Settings.h" from bsdf/compat.h:5 (stale path); forward to the real header
#pragma once
#include <rfw/context/settings.h>
