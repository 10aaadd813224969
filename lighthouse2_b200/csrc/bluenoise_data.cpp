/* bluenoise_data.cpp - embeds data/heitz_bluenoise_256spp.bin (see tools/extract_bluenoise.py) so the
   shared library is self-contained. Layout: sob[65536] | scr[131072] | rnk[131072] bytes. */
#ifndef LH2B_BLUENOISE_PATH
#error "LH2B_BLUENOISE_PATH must point at heitz_bluenoise_256spp.bin"
#endif
__asm__( ".section .rodata\n"
	".global lh2b_bluenoise_bytes\n"
	".hidden lh2b_bluenoise_bytes\n"
	".balign 16\n"
	"lh2b_bluenoise_bytes:\n"
	".incbin \"" LH2B_BLUENOISE_PATH "\"\n"
	".previous\n" );
