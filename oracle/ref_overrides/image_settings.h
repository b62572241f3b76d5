/* Oracle build configuration: shadows the reference's constants/image_settings.h (found first on
 * the include path) so the reference sources compile at the BASELINE.json resolution and method.
 * Same macro names the reference reads; values are ours. Test infrastructure only. */
#ifndef IMAGE_SETTINGS_H
#define IMAGE_SETTINGS_H
#ifndef RLPT_ORACLE_W
#define RLPT_ORACLE_W 512
#endif
#ifndef RLPT_ORACLE_H
#define RLPT_ORACLE_H 512
#endif
#define FULLSCREEN_MODE false
#define SCREEN_WIDTH RLPT_ORACLE_W
#define SCREEN_HEIGHT RLPT_ORACLE_H
#define FOCAL_LENGTH SCREEN_HEIGHT
#define EPS 0.00001f
#define RHO (1.f / (2.f*3.1415926535f))
#define RENDER_SAVED_RADIANCE_VOLUMES false
#define PATH_TRACING_METHOD 1
#endif
