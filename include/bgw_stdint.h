/* Fixed-width integer types: <stdint.h>, except under NVRTC (run-time compilation of the step kernel for one spec,
 * bgw_specialize), which has no host headers. */
#ifndef BGW_STDINT_H
#define BGW_STDINT_H
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long uintptr_t;
#else
#include <stdint.h>
#endif
#endif
