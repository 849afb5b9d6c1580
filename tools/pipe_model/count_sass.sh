#!/bin/bash
# SASS opcode counts inside the loop body of every k_mix instantiation (the loop is the part between the first and last IMAD.WIDE)
cuobjdump -sass "$1" | awk '/Function/{name=$3} /IMAD.WIDE|IADD3|LOP3|SHF|PRMT|IMAD /{for(i=1;i<=NF;i++) if($i ~ /^(IMAD|IADD3|LOP3|SHF|PRMT)/){split($i,p,"."); op=p[1]; if($i ~ /WIDE/) op="WIDE"; if($i ~ /IADD3.X/) op="IADD3.X"; if ($i ~ /IMAD.MOV/) op="IMAD.MOV"; c[name" "op]++}} END{for(k in c) print k, c[k]}' | sort
