#pragma once
#include <any>
namespace boost { using std::any; using std::any_cast; }
