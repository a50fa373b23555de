#pragma once
// Drop-in include: `#include <library/msufsort.h>` with include root `src/`, exactly as the
// reference's demo does (/root/reference/src/executable/msufsort/main.cpp:8).
#include "./msufsort/msufsort.h"
