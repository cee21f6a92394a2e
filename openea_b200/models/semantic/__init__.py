from openea_b200.models._stubs import out_of_scope

DistMult = out_of_scope("DistMult", "bilinear h∘r∘t score: listed as a 'next' variant of kernel K1 (SURVEY §8f)")
HolE = out_of_scope("HolE", "FFT circular correlation")
SimplE = out_of_scope("SimplE", "listed as a 'next' variant of kernel K1 (SURVEY §8f)")
RotatE = out_of_scope("RotatE", "float64 complex rotation")
