#pragma once
#include <memory>
namespace boost { using std::shared_ptr; using std::weak_ptr; using std::dynamic_pointer_cast; using std::static_pointer_cast; }
