/* posix_compat.h - TEST INFRASTRUCTURE ONLY. Force-included (-include) when oracle/Makefile compiles the reference's
   RenderSystem / platform sources on Linux: the handful of Win32 names those sources use outside their own #ifdef WIN32
   blocks (lib/platform/platform.h:117,140-141; lib/RenderSystem/host_scene.cpp:301,332). */
#pragma once
#ifdef __cplusplus
#include <stdio.h>
#include <stddef.h>
typedef void* HANDLE;
struct CRITICAL_SECTION { int unused; };
template <size_t N, class... A> inline int sprintf_s( char (&buf)[N], const char* fmt, A... a ) { return snprintf( buf, N, fmt, a... ); }
#endif
