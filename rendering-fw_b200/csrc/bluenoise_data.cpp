// bluenoise_data.cpp — embeds rendering-fw_b200/data/bluenoise_256spp.bin (tools/extract_bluenoise.py)
// into the library so the C ABI has no run-time file dependency.
#ifndef BLUENOISE_PATH
#error "BLUENOISE_PATH must point at bluenoise_256spp.bin"
#endif
__asm__(".section .rodata\n"
		".balign 16\n"
		".global rfwb200_bluenoise_table\n"
		".hidden rfwb200_bluenoise_table\n"
		"rfwb200_bluenoise_table:\n"
		".incbin \"" BLUENOISE_PATH "\"\n"
		".previous\n");
extern "C" __attribute__((visibility("hidden"))) const unsigned int rfwb200_bluenoise_table_size = 327680;
