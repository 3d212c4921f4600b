#pragma once
#include "../access.hpp"
