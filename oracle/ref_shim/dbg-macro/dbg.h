// stand-in for dbg-macro (develop-build tracing of the reference): evaluates to its last argument. TEST INFRASTRUCTURE ONLY.
#pragma once
#define dbg(...) (__VA_ARGS__)
