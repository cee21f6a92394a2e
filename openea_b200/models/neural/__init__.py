from openea_b200.models._stubs import out_of_scope

ConvE = out_of_scope("ConvE", "2-D convolution scorer")
ProjE = out_of_scope("ProjE", "projection network scorer")
