#pragma once
#include <memory>
namespace boost { using std::enable_shared_from_this; }
