/* oracle/zero_malloc.h -- TEST INFRASTRUCTURE.
 * Force-included (gcc -include) when compiling the UNMODIFIED reference source
 * /root/reference/c/kmers.c into oracle/_ref/kmers_ref_zinit.so, so that its
 * scratch buffers start zeroed.  This pins the one place where the reference
 * reads uninitialised memory (c/kmers.c:279 with :78-91, quirk Q3-UB) to the
 * contract stated in DESIGN.md: never-written reverse-complement slots read
 * as byte 0.  <stdlib.h> is included first so its own declaration of malloc
 * is seen before the macro exists. */
#include <stdlib.h>
#define malloc(n) calloc(1, (n))
