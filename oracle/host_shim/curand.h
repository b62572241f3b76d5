/* oracle/host_shim/curand.h -- TEST INFRASTRUCTURE ONLY: host build of the reference needs nothing from the cuRAND host API. */
#ifndef RLPT_ORACLE_CURAND_HOST_SHIM_H
#define RLPT_ORACLE_CURAND_HOST_SHIM_H
#endif
