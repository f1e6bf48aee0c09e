// TEST INFRASTRUCTURE -- see tf.h
#pragma once
#include "tf.h"
