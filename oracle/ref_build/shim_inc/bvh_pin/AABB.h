// the reference's node headers include "AABB.h"; the file on disk is aabb.h
#pragma once
#include <bvh/aabb.h>
