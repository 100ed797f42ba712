// oracle/shims/pre.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Force-included (-include) ahead of the reference translation unit so that the unmodified
// Windows/VS2010 sources under /root/reference compile with g++ on Linux.
// Nothing here restates reference code; it only supplies what <windows.h>/MSVC's CRT gave it.
#pragma once
#include <cstdlib>
#include <stdlib.h>
#include <math.h>
#include <cmath>
#include <cfloat>
#include <algorithm>
#include <vector>
#include <deque>
#include <string>
#include <iostream>
#include <fstream>
#include <sstream>
#include <iomanip>
#include <limits>

// windows.h supplied min/max as macros; the reference calls them unqualified (MAIN:903,907).
using std::max;
using std::min;

// MSVC spelling used by the matrix header (MAT:715-717).
#define _hypot ::hypot

// POSIX declares random(); the reference declares its own (SIM2:16, MAIN:320).
#define random lqr_ref_random

// MSVC CRT generator: RAND_MAX 32767, LCG 214013/2531011.  The reference's random() assumes
// RAND_MAX*(RAND_MAX+2) fits in int (MAIN:321); with glibc's 2^31-1 it overflows and normal()
// never terminates.
extern "C" int lqr_msvc_rand(void);
extern "C" void lqr_msvc_srand(unsigned int seed);
#undef RAND_MAX
#define RAND_MAX 32767
#define rand lqr_msvc_rand

// The hull is obtained by spawning qconvex.exe (MAIN:769-770,814-815); route the spawn to the
// in-process stand-in in ref_harness.cpp.
extern "C" int lqr_ref_system(const char* cmd);
#define system lqr_ref_system
