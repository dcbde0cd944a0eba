// TEST INFRASTRUCTURE (oracle): see core_min.hpp
#pragma once
#include "core_min.hpp"
