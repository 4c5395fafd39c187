/* TEST INFRASTRUCTURE ONLY.  C++-safe declarations of the handful of wav.c
 * reader entry points the harness uses (wav.h itself is not valid C++: it uses
 * `restrict`; see /root/reference/receiver/wav.h:129-141,212-216). */
#ifndef NAVTEX_ORACLE_WAV_C_API_H
#define NAVTEX_ORACLE_WAV_C_API_H
#include <stddef.h>
#include <stdint.h>
typedef struct _WavFile WavFile;
#define WAV_OPEN_READ 1
WavFile *wav_open(const char *filename, uint32_t mode);
void wav_close(WavFile *self);
size_t wav_read(WavFile *self, void *buffer, size_t count);
uint16_t wav_get_num_channels(const WavFile *self);
size_t wav_get_sample_size(const WavFile *self);
uint32_t wav_get_sample_rate(const WavFile *self);
#endif
