/*
 * map_file.h -- read-only file mapping for the host tools.
 * Same interface as the reference's file mapper (/root/reference/map_file.h:8-16):
 * map_input_file2(), map_input_file(), unmap_file() and the map_in() macro, so the
 * three mains keep the reference's input path unchanged.
 */
#ifndef BWTS_B200_MAP_FILE_H
#define BWTS_B200_MAP_FILE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	void *sp, *ep;   /* first byte, one past the last byte */
} ptr_range;

/* Map `filename` read-only and private.  On any failure: perror() and exit(1), like the
 * reference (a zero-length file fails in mmap with "Invalid argument"). */
void map_input_file2(const char *filename, void **start, long *len);
ptr_range map_input_file(const char *filename);
void unmap_file(ptr_range extent);

#define map_in(ptr, len, path) map_input_file2(path, (void **)&ptr, &len), len /= sizeof(*ptr)

#ifdef __cplusplus
}
#endif

#endif
