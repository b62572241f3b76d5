/* Baseline build configuration: shadows C/constants/image_settings.h (found first on the include path). Same macro names and
 * values as committed (512 x 512, FOCAL_LENGTH = SCREEN_HEIGHT) except PATH_TRACING_METHOD 3 = the default path tracer
 * (BASELINE.json configs[0]; the committed value is 0). */
#ifndef IMAGE_SETTINGS_H
#define IMAGE_SETTINGS_H
#define FULLSCREEN_MODE false
#ifndef RLPT_CE_W
#define RLPT_CE_W 512
#endif
#ifndef RLPT_CE_H
#define RLPT_CE_H 512
#endif
#define SCREEN_WIDTH RLPT_CE_W
#define SCREEN_HEIGHT RLPT_CE_H
#define FOCAL_LENGTH SCREEN_HEIGHT
#define EPS 0.00001f
#define PATH_TRACING_METHOD 3
#endif
