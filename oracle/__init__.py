"""TEST INFRASTRUCTURE ONLY -- see oracle/muscl_oracle.h.  Never imported by the product package."""
