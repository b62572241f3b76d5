/* Stub standing in for SDL2's SDL.h so the reference engine's headers parse without SDL2
 * installed. The oracle harness never opens a window. Test infrastructure only. */
#ifndef RLPT_ORACLE_SDL_STUB_H
#define RLPT_ORACLE_SDL_STUB_H
struct SDL_Window; struct SDL_Renderer; struct SDL_Texture;
#endif
