// RcppEigen.h -- the hot-path natives include it but use no Eigen type (only mvnorm.cpp does, which is NOT built: Eigen is absent).
#pragma once
#include "Rcpp.h"
