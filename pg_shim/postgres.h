/*
 * pg_shim/postgres.h -- minimal stand-in for PostgreSQL's postgres.h.
 *
 * There is no PostgreSQL source tree in this image.  This shim declares only
 * what the reference's compression.c / storage.c (and our drop-in
 * host/compression.c) use, so those files compile unchanged outside a server.
 * It is build scaffolding, not product code.
 */
#ifndef PG_SHIM_POSTGRES_H
#define PG_SHIM_POSTGRES_H

#include <assert.h>
#include <setjmp.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef size_t Size;
typedef uint8_t uint8;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef uint64_t uint64;
typedef int32_t int32;
typedef uint32 TransactionId;
typedef uint32 BlockNumber;
typedef uint16 LocationIndex;
typedef unsigned int Oid;

typedef struct
{
    uint32 xlogid;
    uint32 xrecoff;
} PageXLogRecPtr;

#define BLCKSZ 8192
#define InvalidBlockNumber ((BlockNumber) 0xFFFFFFFF)

#define DEBUG1 14
#define ERROR 20

/*
 * elog(ERROR) in PostgreSQL longjmps to the error handler.  The shim does the
 * same when a test has armed pg_shim_error_jmp, otherwise it aborts.
 */
extern sigjmp_buf *pg_shim_error_jmp;
extern char pg_shim_last_error[256];
void pg_shim_elog(int level, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
#define elog(level, ...) pg_shim_elog((level), __VA_ARGS__)

#define palloc(sz) malloc(sz)
#define pfree(p) free(p)
#define Assert(c) ((void) 0)
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))

#endif
