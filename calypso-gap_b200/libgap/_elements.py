"""Chemical symbol -> atomic number, Z = 0 ('X') .. 103 ('Lr'), the range the
reference's lookup table covers (gappy/libgap/GAP.py:21-31)."""
_SYMBOLS = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn "
    "Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce "
    "Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn "
    "Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr"
).split()
ATOMIC_NUMBER = {s: z for z, s in enumerate(_SYMBOLS)}


def atomic_numbers(symbols):
    """KeyError / TypeError if any entry is not a known symbol (callers fall back
    to treating the input as atomic numbers, like the reference's try/except)."""
    return [ATOMIC_NUMBER[s] for s in symbols]
