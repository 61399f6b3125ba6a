/* pg_shim/access/htup.h -- HeapTupleData as storage.c sees it. */
#ifndef PG_SHIM_HTUP_H
#define PG_SHIM_HTUP_H
#include "postgres.h"

typedef struct
{
    uint16 bi_hi;
    uint16 bi_lo;
    uint16 ip_posid;
} ItemPointerData;

typedef struct HeapTupleHeaderData *HeapTupleHeader;

typedef struct HeapTupleData
{
    uint32          t_len;
    ItemPointerData t_self;
    Oid             t_tableOid;
    HeapTupleHeader t_data;
} HeapTupleData;
typedef HeapTupleData *HeapTuple;
#endif
