/* pg_shim/utils/guc.h -- the three GUC-definition entry points compression.c uses. */
#ifndef PG_SHIM_GUC_H
#define PG_SHIM_GUC_H
#include "postgres.h"

struct config_enum_entry
{
    const char *name;
    int         val;
    bool        hidden;
};

#define PGC_USERSET 6

void DefineCustomEnumVariable(const char *name, const char *short_desc, const char *long_desc,
                              int *valueAddr, int bootValue,
                              const struct config_enum_entry *options, int context, int flags,
                              void *check_hook, void *assign_hook, void *show_hook);
void DefineCustomIntVariable(const char *name, const char *short_desc, const char *long_desc,
                             int *valueAddr, int bootValue, int minValue, int maxValue,
                             int context, int flags,
                             void *check_hook, void *assign_hook, void *show_hook);
#endif
