// main.cpp -- `hifimeth-b200 call [OPTIONS] BAM MOD-BAM`: the sub-command table of the reference
// (src/app/hifimeth/main.cpp:35-58) reduced to the one command this engine replaces.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <unistd.h>

extern "C" int hm_call_main(int argc, char** argv);
extern "C" void hm_call_fast_exit(int on);

int main(int argc, char** argv)
{
    if (argc >= 2 && strcmp(argv[1], "call") == 0) {
        // the output file is closed when hm_call_main returns; the engines' teardown (pinned memory, CUDA contexts: up to seconds on
        // a multi-GPU box) is left to the driver's process-exit clean-up
        hm_call_fast_exit(1);
        const int rc = hm_call_main(argc, argv);
        fflush(nullptr);
        _exit(rc);
    }
    fprintf(stderr, "USAGE:\n  %s call [OPTIONS] BAM MOD-BAM\n\nOnly the `call` command of hifimeth is provided by the B200 engine.\n", argc ? argv[0] : "hifimeth-b200");
    return EXIT_FAILURE;
}
