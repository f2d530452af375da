// empty stub: the reference includes this boost header (AlignmentGraph.cpp:4-7) but never uses it
