// Frontend path (BRISK-style detect/describe + Hamming matching) -- implemented in a later commit.
#include "okb_ctx.h"
void okb_frontend_release(okb_ctx*) {}
