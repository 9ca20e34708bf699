// TEST INFRASTRUCTURE ONLY: openvdb/tools/VelocityFields.h:232 uses BOOST_STATIC_ASSERT without
// including a Boost header that defines it (it leaks in through other Boost headers in a real install).
#pragma once
#ifndef BOOST_STATIC_ASSERT
#define BOOST_STATIC_ASSERT(...) static_assert(__VA_ARGS__, #__VA_ARGS__)
#endif
