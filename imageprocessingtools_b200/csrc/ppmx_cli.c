/* ppmx_cli.c -- the command line of ppmx-edward (ref:117-191 of /root/reference/ppmx-edward.c),
 * same flags, same "<input>.out" result, pixel loops on the GPU.  See include/ppmx_host.h. */
#include "../../include/ppmx_host.h"

int main(int argc, char *argv[]) { return ppmx_main(argc, argv); }
