/* oracle/ref_host_radiance_map.cpp -- TEST INFRASTRUCTURE ONLY.
 * Host (g++) build of the reference's radiance_volumes/radiance_map.cu, included from where it lies under
 * /root/reference. g++ rejects one construct that nvcc's front end tolerates: in
 * RadianceMap::temporal_difference_update_radiance_volume_sector (radiance_map.cu:113-145) `case SURFACE:`
 * jumps past the initialised local declared under `case AREA_LIGHT:`. Every case of that switch ends in
 * exactly one `break;`, and it is the only switch/break in the file, so wrapping each case in its own block
 * with two macros makes it well-formed without touching the reference source. All headers are included
 * first so the macros only ever see radiance_map.cu's own text. */
#include "radiance_map.cuh"
#include <algorithm>
#include <iostream>
#include <ctime>
#include "printing.h"
#define case { case
#define break break; }
#include "radiance_map.cu"
#undef case
#undef break
