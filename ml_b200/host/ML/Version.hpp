#pragma once
// Same version string as the reference this is a drop-in for (ML/Version.hpp:7), plus a backend tag.
#define MLPP_VERSION "0.8.0"
#define MLPP_BACKEND "b200-cuda"
