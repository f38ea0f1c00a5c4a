/*
 * Minimal stand-in for torch-0.4's <TH/TH.h>, written for this repo.
 *
 * TEST INFRASTRUCTURE ONLY.  It exists so that the reference's own, unmodified
 * C sources (roialign/roi_align/src/crop_and_resize.c and nms/src/nms.c under
 * /root/reference) can be compiled into oracle/_ref/ as the parity checker and
 * CPU baseline.  Nothing in the product path includes this file.
 *
 * Only the handful of TH entry points those two files touch are provided:
 *   - tensors are {data, size[4], nd}; crop_and_resize.c reads ->size[i]
 *     directly (crop_and_resize.c:124-129,165-172), so `size` is a member array;
 *   - resize4d only records the sizes (the caller pre-allocates the storage);
 *   - zero is a memset over prod(size);
 *   - the byte tensor used as nms.c's `suppressed` scratch is malloc/free.
 */
#ifndef SLN_ORACLE_TH_SHIM_H
#define SLN_ORACLE_TH_SHIM_H

#include <stdlib.h>
#include <string.h>

#define SLN_TH_DECL(NAME, T)                                   \
    typedef struct NAME {                                      \
        T *data;                                               \
        long size[4];                                          \
        int nd;                                                \
    } NAME;                                                    \
    static inline T *NAME##_data(NAME *t) { return t->data; }  \
    static inline long NAME##_size(NAME *t, int d) { return t->size[d]; }

SLN_TH_DECL(THFloatTensor, float)
SLN_TH_DECL(THIntTensor, int)
SLN_TH_DECL(THLongTensor, long)
SLN_TH_DECL(THByteTensor, unsigned char)

static inline long sln_th_numel(const long *size, int nd)
{
    long n = 1;
    for (int i = 0; i < nd; ++i) n *= size[i];
    return n;
}

static inline void THFloatTensor_resize4d(THFloatTensor *t, long a, long b, long c, long d)
{
    t->size[0] = a; t->size[1] = b; t->size[2] = c; t->size[3] = d;
    t->nd = 4;
}

static inline void THFloatTensor_zero(THFloatTensor *t)
{
    memset(t->data, 0, sizeof(float) * (size_t)sln_th_numel(t->size, t->nd));
}

static inline THByteTensor *THByteTensor_newWithSize1d(long n)
{
    THByteTensor *t = (THByteTensor *)malloc(sizeof(THByteTensor));
    t->data = (unsigned char *)malloc((size_t)(n > 0 ? n : 1));
    t->size[0] = n; t->size[1] = t->size[2] = t->size[3] = 1;
    t->nd = 1;
    return t;
}

static inline void THByteTensor_fill(THByteTensor *t, unsigned char v)
{
    memset(t->data, v, (size_t)t->size[0]);
}

static inline void THByteTensor_free(THByteTensor *t)
{
    free(t->data);
    free(t);
}

#define THLongTensor_isContiguous(x) 1
#define THArgCheck(cond, argn, msg) do { (void)(cond); } while (0)

#endif
