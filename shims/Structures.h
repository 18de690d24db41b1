#pragma once
#include <rfw/context/structs.h>
