#pragma once
#include <memory>
namespace boost { using std::shared_ptr; }
