#pragma once
#include "highgui.hpp"
