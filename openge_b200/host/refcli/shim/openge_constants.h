// Generated-header stand-in (the reference renders openge_constants.h.in with CMake).
#define OPENGE_VERSION_STRING "0.3-dev"
#define OPENGE_VERSION "0.3"
#define OPENGE_BUILD_TYPE "dev"
