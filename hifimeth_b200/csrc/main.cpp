// main.cpp -- `hifimeth-b200 call [OPTIONS] BAM MOD-BAM`: the sub-command table of the reference
// (src/app/hifimeth/main.cpp:35-58) reduced to the one command this engine replaces.
#include <cstdio>
#include <cstdlib>
#include <cstring>

extern "C" int hm_call_main(int argc, char** argv);

int main(int argc, char** argv)
{
    if (argc >= 2 && strcmp(argv[1], "call") == 0) return hm_call_main(argc, argv);
    fprintf(stderr, "USAGE:\n  %s call [OPTIONS] BAM MOD-BAM\n\nOnly the `call` command of hifimeth is provided by the B200 engine.\n", argc ? argv[0] : "hifimeth-b200");
    return EXIT_FAILURE;
}
