/* pg_shim/access/htup_details.h -- ItemId, MAXALIGN, MaxHeapTuplesPerPage. */
#ifndef PG_SHIM_HTUP_DETAILS_H
#define PG_SHIM_HTUP_DETAILS_H
#include "access/htup.h"

typedef struct ItemIdData
{
    unsigned lp_off:15, lp_flags:2, lp_len:15;
} ItemIdData;
typedef ItemIdData *ItemId;

#define MAXALIGN(LEN) (((uintptr_t) (LEN) + 7) & ~((uintptr_t) 7))
#define SizeOfPageHeaderData 24
#define SizeofHeapTupleHeader 23
/* (BLCKSZ - SizeOfPageHeaderData) / (MAXALIGN(SizeofHeapTupleHeader) + sizeof(ItemIdData)) */
#define MaxHeapTuplesPerPage ((int) ((BLCKSZ - SizeOfPageHeaderData) / (24 + 4)))
#endif
