/* Oracle build configuration: shadows the reference's constants/monte_carlo_settings.h.
 * Same macro names; spp and environment light can be chosen per oracle build. */
#ifndef MONTE_CARLO_SETTING_H
#define MONTE_CARLO_SETTING_H
#ifndef RLPT_ORACLE_SPP
#define RLPT_ORACLE_SPP 32
#endif
#ifndef RLPT_ORACLE_ENV
#define RLPT_ORACLE_ENV 0.0f
#endif
#ifndef RLPT_ORACLE_BOUNCES
#define RLPT_ORACLE_BOUNCES 80
#endif
#define MAX_RAY_BOUNCES RLPT_ORACLE_BOUNCES
#define SAMPLES_PER_PIXEL RLPT_ORACLE_SPP
#define ENVIRONMENT_LIGHT RLPT_ORACLE_ENV
#define THROUGHPUT_THRESHOLD 0.0001f
#endif
