/* pg_shim/pg_shim.c -- runtime part of the PostgreSQL header shim. */
#include "postgres.h"
#include "utils/guc.h"
#include <stdarg.h>

sigjmp_buf *pg_shim_error_jmp = NULL;
char pg_shim_last_error[256];

void
pg_shim_elog(int level, const char *fmt, ...)
{
    va_list ap;

    va_start(ap, fmt);
    vsnprintf(pg_shim_last_error, sizeof(pg_shim_last_error), fmt, ap);
    va_end(ap);
    if (level < ERROR)
        return;
    if (pg_shim_error_jmp)
        siglongjmp(*pg_shim_error_jmp, 1);
    fprintf(stderr, "ERROR: %s\n", pg_shim_last_error);
    abort();
}

void
DefineCustomEnumVariable(const char *name, const char *short_desc, const char *long_desc,
                         int *valueAddr, int bootValue,
                         const struct config_enum_entry *options, int context, int flags,
                         void *check_hook, void *assign_hook, void *show_hook)
{
    (void) name; (void) short_desc; (void) long_desc; (void) options; (void) context;
    (void) flags; (void) check_hook; (void) assign_hook; (void) show_hook;
    *valueAddr = bootValue;
}

void
DefineCustomIntVariable(const char *name, const char *short_desc, const char *long_desc,
                        int *valueAddr, int bootValue, int minValue, int maxValue,
                        int context, int flags,
                        void *check_hook, void *assign_hook, void *show_hook)
{
    (void) name; (void) short_desc; (void) long_desc; (void) minValue; (void) maxValue;
    (void) context; (void) flags; (void) check_hook; (void) assign_hook; (void) show_hook;
    *valueAddr = bootValue;
}
