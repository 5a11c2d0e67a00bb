function [Floe, kill, transfer, info] = sz_resident_timestep(Floe, floebound, ocean, winds, c2_boundary, dt, HFo, Nb, doInt, COLLISION, PERIODIC, Modulus, topology_changed, want_rows)
%SZ_RESIDENT_TIMESTEP  One SubZero timestep with the floe state resident on the GPU: the contact loop of
% floe_interactions_all.m:16-277, the ocean/atmosphere forcing and the integrator of calc_trajectory.m (:281) run on the
% device; MATLAB receives the per-floe state every step and the contact rows only when asked (corners every 10 steps,
% fracture every 75: Subzero.m:317-352).  Ridging, rafting and the kill/fuse tail (floe_interactions_all.m:288-512) stay
% the reference's code and run on the returned struct array.
%
%   topology_changed  true on the first call and after anything that created, removed or reshaped floes (fracture, weld,
%                     ridge, raft, FloeSimplify, corners, new ice): the whole state is uploaded again.
%   want_rows         true when Floe(i).interactions is needed on the MATLAB side this step.
%
% The ocean grid and the Monte-Carlo points (Floe.X/.Y/.A, initialize_floe_values.m:31-33) travel with the upload.
    persistent voff
    N0 = numel(Floe);
    tp = struct('dt', dt, 'HFo', mean(HFo(:)), 'xo_min', min(ocean.Xo), 'xo_max', max(ocean.Xo), 'yo_min', min(ocean.Yo), 'yo_max', max(ocean.Yo));
    if topology_changed
        prm = struct('Lx', max(c2_boundary(1,:)), 'Ly', max(c2_boundary(2,:)), 'modulus', Modulus, 'dt', dt, ...
                     'Nb', Nb, 'periodic', double(PERIODIC), 'collision', double(COLLISION));
        nv = arrayfun(@(f) size(f.c_alpha, 2), Floe);
        ca = [Floe.c_alpha];  c0 = [Floe.c0];
        voff = [0; cumsum(nv(:))];
        soa = struct('x', cat(1, Floe.Xi), 'y', cat(1, Floe.Yi), 'rmax', cat(1, Floe.rmax), 'h', cat(1, Floe.h), ...
                     'area', cat(1, Floe.area), 'u', cat(1, Floe.Ui), 'v', cat(1, Floe.Vi), 'ksi', cat(1, Floe.ksi_ice), ...
                     'alive', double(cat(1, Floe.alive)), 'voff', voff, 'vx', ca(1,:)', 'vy', ca(2,:)');
        if PERIODIC
            sz_resident_mex('upload', prm, soa);
        else
            hv = holes(floebound.poly).Vertices;                 % floe_interactions.m:31
            bnd = struct('x', hv(:,1), 'y', hv(:,2), 'box_x', c2_boundary(1,:)', 'box_y', c2_boundary(2,:)', 'area', floebound.area, 'h', floebound.h, ...
                         'xi', floebound.Xi, 'yi', floebound.Yi, 'u', floebound.Ui, 'v', floebound.Vi, 'ksi', floebound.ksi_ice);
            sz_resident_mex('upload', prm, soa, bnd);
        end
        z = @(name) cat(1, Floe.(name));
        st = struct('mass', z('mass'), 'inertia', z('inertia_moment'), 'alpha', z('alpha_i'), 'dXi_p', z('dXi_p'), 'dYi_p', z('dYi_p'), ...
                    'dUi_p', z('dUi_p'), 'dVi_p', z('dVi_p'), 'dalpha_p', z('dalpha_i_p'), 'dksi_p', z('dksi_ice_p'), ...
                    'FxOA', z('FxOA'), 'FyOA', z('FyOA'), 'torqueOA', z('torqueOA'), 'c0x', c0(1,:)', 'c0y', c0(2,:)', 'nz', size(Floe(1).StressH, 3));
        sz_resident_mex('trajectory_init', st);
        sz_resident_mex('set_ocean', ocean, winds);
        sz_resident_mex('set_points', [Floe.X], [Floe.Y], double([Floe.A]));
    end
    info = sz_resident_mex('step');                               % floe_interactions_all.m:16-277
    out = sz_resident_mex('floe_outputs');
    kill = out.kill';  transfer = out.transfer';
    if want_rows
        r = sz_resident_mex('rows');
        for i = 1+Nb:N0
            Floe(i).interactions = r.rows(:, r.row_off(i)+1 : r.row_off(i+1))';
        end
    end
    sz_resident_mex('ocean_forcing', tp, double(doInt.flag));     % calc_trajectory.m:94-166 (doInt.flag, or floes thinner than 0.1 m)
    info.sacked = sz_resident_mex('trajectory_step', tp);         % calc_trajectory.m:3-46,67-80,170-234
    s = sz_resident_mex('state');
    for i = 1+Nb:N0
        Floe(i).OverlapArea = out.overlap_area(i);
        Floe(i).collision_force = [out.fx(i) out.fy(i)];
        Floe(i).collision_torque = out.torque(i);
        Floe(i).potentialInteractions = [];
        if bitand(s.flags(i), 1), kill(i) = i; continue; end      % sacked: the caller keeps the old struct (:282)
        Floe(i).Xi = s.x(i);  Floe(i).Yi = s.y(i);  Floe(i).Ui = s.u(i);  Floe(i).Vi = s.v(i);  Floe(i).ksi_ice = s.ksi(i);
        Floe(i).h = s.h(i);  Floe(i).alive = s.alive(i);  Floe(i).mass = s.mass(i);  Floe(i).inertia_moment = s.inertia(i);
        Floe(i).alpha_i = s.alpha(i);  Floe(i).dXi_p = s.dXi_p(i);  Floe(i).dYi_p = s.dYi_p(i);  Floe(i).dUi_p = s.dUi_p(i);  Floe(i).dVi_p = s.dVi_p(i);
        Floe(i).dalpha_i_p = s.dalpha_p(i);  Floe(i).dksi_ice_p = s.dksi_p(i);
        Floe(i).Stress = reshape(s.stress(:, i), 2, 2)';  Floe(i).strain = reshape(s.strain(:, i), 2, 2)';
        Floe(i).FxOA = s.FxOA(i);  Floe(i).FyOA = s.FyOA(i);  Floe(i).torqueOA = s.torqueOA(i);
        k = voff(i)+1 : voff(i+1);
        Floe(i).c_alpha = [s.cax(k)'; s.cay(k)'];                  % A_rot * c0 (calc_trajectory.m:221-222)
    end
end
