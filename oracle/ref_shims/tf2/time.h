#include "../lpref_tf2.hpp"
