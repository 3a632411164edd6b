// tests/ref_sanitize_main.cpp -- TEST INFRASTRUCTURE ONLY.
// main() around harness_ref_map so that tests/test_ref_host.py can run the mecat2ref stage sequence and kernel bodies under
// AddressSanitizer / UndefinedBehaviorSanitizer: every "device" array of the host harness is an exact-size malloc, so an
// index the kernels would get wrong on the GPU (block tables, candidate lists, hit counts) is reported here.
#include <stdio.h>
#include <stdlib.h>
extern "C" int harness_ref_map(const char*, const char*, int, int, int, int, long, char**, size_t*, long*, char*, int);
int main(int argc, char** argv)
{
	char* text; size_t n; long st[4]; char err[512];
	int rc = harness_ref_map(argv[1], argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atol(argv[7]), &text, &n, st, err, 512);
	printf("rc=%d bytes=%zu tasks=%ld batches=%ld rescue=%ld with_strings=%ld %s\n", rc, n, st[0], st[1], st[2], st[3], rc ? err : "");
	free(text);
	return rc;
}
