/* Baseline build configuration: shadows C/constants/monte_carlo_settings.h. Committed values: MAX_RAY_BOUNCES 2, SAMPLES_PER_PIXEL 16;
 * a second build uses 80 bounces, the GPU engine's setting (G/constants/monte_carlo_settings.h:8). */
#ifndef MONTE_CARLO_SETTING_H
#define MONTE_CARLO_SETTING_H
#ifndef RLPT_CE_BOUNCES
#define RLPT_CE_BOUNCES 2
#endif
#ifndef RLPT_CE_SPP
#define RLPT_CE_SPP 16
#endif
#define MAX_RAY_BOUNCES RLPT_CE_BOUNCES
#define SAMPLES_PER_PIXEL RLPT_CE_SPP
#endif
