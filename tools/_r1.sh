timeout 300 python -m pytest tests -m gpu -x -q -k "route" 2>&1 | tail -2
cd tools
timeout 120 python ab.py --route 2>&1 | grep "route W"
DINT_ROUTE_GRID=592 timeout 120 python ab.py --route 2>&1 | grep "route W"
DINT_ROUTE_GRID=296 timeout 120 python ab.py --route 2>&1 | grep "route W"
SANITY_MODES=p2p timeout 100 python p2p_sanity.py 2>&1 | grep "p2p:"
