// Rcpp::SparseMatrix lives in the stub Rcpp.h. TEST INFRASTRUCTURE ONLY.
#pragma once
