#pragma once
#include "_std_thread.hpp"
