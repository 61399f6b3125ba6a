/* pg_shim/storage/bufpage.h -- nothing from bufpage.h is used by the codec path. */
