/* Stub standing in for SDL2's SDL.h (SDL2 is not installed): the CPU-engine baseline never opens a window; its SDLScreen is the
 * headless one in cpu_engine_harness.cpp. Baseline infrastructure only. */
#ifndef RLPT_CPU_ENGINE_SDL_STUB_H
#define RLPT_CPU_ENGINE_SDL_STUB_H
struct SDL_Window; struct SDL_Renderer; struct SDL_Texture;
#endif
