/* ORACLE / TEST INFRASTRUCTURE.  Stand-in for <fftw3.h>: gri_fft_filter_ccc_generic.cc includes it but only uses
 * gri_fft_complex, which oracle/shim/gri_fft.h implements (FFTW 3 itself is third party and absent). */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H
#endif
