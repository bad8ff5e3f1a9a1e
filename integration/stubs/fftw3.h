/* stub of <fftw3.h> for `make refcheck` */
#pragma once
typedef float fftwf_complex[2];
typedef struct fftwf_plan_s* fftwf_plan;
