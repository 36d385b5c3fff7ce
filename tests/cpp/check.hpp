// minimal REQUIRE / REQUIRE_THROWS_AS (Catch2 is fetched from the network by the reference's test build
// and is not available here)
#pragma once
#include <cstdio>
#include <cstdlib>
#define REQUIRE(cond)                                                                 \
    do {                                                                              \
        if (!(cond)) {                                                                \
            std::fprintf(stderr, "REQUIRE failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            std::exit(1);                                                             \
        }                                                                             \
    } while (0)
#define REQUIRE_THROWS_AS(expr, type)                                                 \
    do {                                                                              \
        bool thrown_ = false;                                                         \
        try { (void) (expr); } catch (const type &) { thrown_ = true; } catch (...) {} \
        if (!thrown_) {                                                               \
            std::fprintf(stderr, "REQUIRE_THROWS_AS failed: %s (%s:%d)\n", #expr, __FILE__, __LINE__); \
            std::exit(1);                                                             \
        }                                                                             \
    } while (0)
