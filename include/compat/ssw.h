/* ssw.h compatibility shim: lets the reference's indelpost/sswpy.pyx (`cdef extern from "ssw.h"`,
 * sswpy.pyx:57-83) compile and link against libswb200 unchanged.  See INTEGRATION.md §1. */
#ifndef SWB200_COMPAT_SSW_H
#define SWB200_COMPAT_SSW_H
#include "../swb200.h"
#ifdef __cplusplus
extern "C" {
#endif
static inline char     cigar_int_to_op(uint32_t c)  { return swb_cigar_int_to_op(c); }     /* ssw.h:178-181 */
static inline uint32_t cigar_int_to_len(uint32_t c) { return swb_cigar_int_to_len(c); }    /* ssw.h:187-189 */
static inline uint32_t to_cigar_int(uint32_t l, char o) { return swb_to_cigar_int(l, o); } /* ssw.h:171-173 */
#ifdef __cplusplus
}
#endif
#endif
