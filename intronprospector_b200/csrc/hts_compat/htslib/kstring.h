/* kstring.h -- minimal stand-in for <htslib/kstring.h>; see hts.h in this directory. */
#ifndef IPX_COMPAT_KSTRING_H
#define IPX_COMPAT_KSTRING_H
#include <stddef.h>
typedef struct kstring_t {
    size_t l, m;
    char *s;
} kstring_t;
#endif
